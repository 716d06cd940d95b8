#!/usr/bin/env python3
"""Shared-memory bank model of the FAST kernel's arc-score loads (CPU only): for the pixels that pass the precheck in
every 30-px cell of a synthetic 640x480 frame, queue them in a given order, take them 32 at a time (one warp step of
phase 2) and count, for each of the 17 byte loads, the wavefronts = max over banks of distinct words.  Reproduces the
measured replay rate of the shipped layout (raster order, 11 words per tile row: 1.65 wavefronts per load = 39 % replays;
ncu: 42 %) and shows that the other queue orders tried here are worse and that a 13-word pitch would save 11 % of the
wavefronts (measured: no change in kernel time -- the LSU wavefronts are not the only limiter).
usage: python tools/fast_bank_sim.py"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
synth = importlib.import_module('jetracer-orbslam2_b200.synth')
img = synth.textured_frame(640,480,5000).astype(np.int32)
H,W = img.shape
t=20
offs = [(0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)]
def flagged(cell):
    # cell: (y0,x0,ch,cw) tested range; returns list of (y,x) in-cell coords passing the precheck, raster order
    y0,x0,ch,cw = cell
    c = img[y0:y0+ch, x0:x0+cw]
    def sh(dy,dx): return img[y0+dy:y0+dy+ch, x0+dx:x0+dx+cw]
    an = lambda d: (np.abs(d)>>1) >= ((t+1)>>1)
    f = (an(sh(3,0)-c)|an(sh(-3,0)-c)) & (an(sh(0,3)-c)|an(sh(0,-3)-c))
    ys,xs = np.nonzero(f)
    return ys,xs
def wavefronts(ys,xs,tpw,order):
    idx = order(ys,xs)
    ys,xs = ys[idx],xs[idx]
    tot=0; n=0
    for b in range(0,len(ys),32):
        yy,xx = ys[b:b+32], xs[b:b+32]
        for dx,dy in offs+[(0,0)]:
            byte = (yy+3+dy)*(tpw*4) + xx+4+dx
            word = byte//4
            bank = word%32
            # wavefronts = max over banks of distinct words in that bank
            wf = 0
            for bk in np.unique(bank):
                wf = max(wf, len(np.unique(word[bank==bk])))
            tot+=wf; n+=1
    return tot, n
cells=[(19+30*i, 19+30*j, 30, 30) for i in range(14) for j in range(20)]
orders = {
 'raster (current)': lambda y,x: np.lexsort((x,y)),
 'column-major': lambda y,x: np.lexsort((y,x)),
 'by (y%3, y, x)': lambda y,x: np.lexsort((x,y,y%3)),
 'by (x//4, y)': lambda y,x: np.lexsort((x%4, y, x//4)),
 'by 8x8 tiles': lambda y,x: np.lexsort((x,y,x//8,y//8)),
}
for tpw in (11,13):
    for name,o in orders.items():
        T=N=0
        for c in cells:
            ys,xs = flagged(c)
            if len(ys)==0: continue
            a,b = wavefronts(ys,xs,tpw,o); T+=a; N+=b
        print(f'tile pitch {tpw} words, order {name:20s}: {T/N:.2f} wavefronts per warp load')
