"""Static SASS checks (no GPU): the library contains sm_100a code only, tensor-core instructions only where the path has a
dense contraction (the brute-force Hamming matcher: int8 IMMA), and each kernel carries the instructions its design
leans on -- the evidence profiles/r01m_sass_mnemonics.txt records, kept true by a test."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jetracer-orbslam2_b200", "liborbb200.so")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    import __graft_entry__ as g
    g.build()
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    return kernels, archs


def _count(ops, prefix):
    return sum(1 for o in ops if o == prefix or o.startswith(prefix + "."))


def _kernel(kernels, *needles):
    hits = [k for k in kernels if all(n in k for n in needles)]
    assert hits, needles
    return kernels[hits[0]]


def test_sm100a_only_and_tensor_cores_only_in_the_matcher(sass):
    kernels, archs = sass
    assert archs == {"sm_100a"}
    assert len(kernels) >= 25
    for name, ops in kernels.items():
        for tc in ("HMMA", "DMMA", "UTCHMMA", "UTCQMMA", "WGMMA"):
            assert _count(ops, tc) == 0, (name, tc)
        # the one dense contraction of the path -- all-pairs Hamming distances -- runs as int8 MMAs; nothing else may:
        # warp-level IMMA in k_match_imma, tcgen05 (UTCIMMA, accumulators read back with LDTM) in k_match_umma
        if "k_match_imma" not in name and "k_imma_rate" not in name:
            assert _count(ops, "IMMA") == 0, name
        if "k_match_umma" not in name:
            assert _count(ops, "UTCIMMA") == 0 and _count(ops, "LDTM") == 0, name


def test_kernels_carry_their_instructions(sass):
    kernels, _ = sass
    match1 = _kernel(kernels, "k_matchILi1E")
    assert _count(match1, "POPC") >= 5 and _count(match1, "LOP3") >= 14  # carry-save popcount: 5 POPC per pair
    for kk in ("k_matchILi1E", "k_matchILi2E"):  # query descriptors and running bests stay in registers
        ops = _kernel(kernels, kk)
        assert _count(ops, "LDL") == 0 and _count(ops, "STL") == 0, kk
    for kk in ("k_match_immaILi1E", "k_match_immaILi2E"):  # 32 MMAs per pair of column groups, A fragments in registers
        ops = _kernel(kernels, kk)
        assert _count(ops, "IMMA.16832.S8.S8") >= 32 and _count(ops, "LDS.128") >= 8, kk
        assert _count(ops, "LDL") == 0 and _count(ops, "STL") == 0 and _count(ops, "POPC") == 0, kk
    for kk in ("k_match_ummaILi1ELb0E", "k_match_ummaILi2ELb0E", "k_match_ummaILi1ELb1E", "k_match_ummaILi2ELb1E"):
        # tcgen05: 2 A tiles x 8 K steps per train tile, 128 columns per LDTM round; B tiles expanded by the workers
        # (Lb0: 16-byte stores of the A tiles, the first two and the refill B tiles) or bulk-copied from the image (Lb1)
        ops = _kernel(kernels, kk)
        assert _count(ops, "UTCIMMA") == 16 and _count(ops, "LDTM") == 4, kk
        assert _count(ops, "STS.128") >= (16 if "Lb1E" in kk else 24) and _count(ops, "UBLKCP") == (1 if "Lb1E" in kk else 0), kk
        assert _count(ops, "VIMNMX3") >= 56, kk  # the chunk-maximum pass in front of the packed-key pass
        assert _count(ops, "LDL") == 0 and _count(ops, "STL") == 0 and _count(ops, "POPC") == 0, kk
    fast = _kernel(kernels, "k_fast_cellsILb0ELb0E")
    assert _count(fast, "VABSDIFF4") >= 4 and _count(fast, "VIMNMX3") >= 40  # packed precheck, arc-score min/max network
    assert _count(fast, "ATOMG") + _count(fast, "RED") >= 1                  # cell-table atomics while emitting
    assert _count(fast, "ACQBULK") == 1                                       # programmatic dependent of the pyramid chain
    resize = _kernel(kernels, "k_resize_rowsILb0E")
    assert _count(resize, "IDP.2A") == 16 and _count(resize, "ACQBULK") == 1 and _count(resize, "PREEXIT") == 1
    # the PDL wait precedes every load of the source level
    first_ld = min(i for i, o in enumerate(resize) if o.startswith("LDG"))
    assert resize.index("ACQBULK") < first_ld
    first_ld = min(i for i, o in enumerate(fast) if o.startswith("LDG"))
    assert fast.index("ACQBULK") < first_ld
    blur = _kernel(kernels, "k_blur")
    assert _count(blur, "IDP.4A") >= 8
    octree = _kernel(kernels, "k_octreeILi1ELi4E")
    assert _count(octree, "REDUX") >= 32 and _count(octree, "BAR") >= 1
    assert _count(octree, "LDS") >= 20 and _count(octree, "LDL") == 0 and _count(octree, "STL") == 0  # pyramid in shared memory, 32 registers, no spills
    cta = _kernel(kernels, "k_fast_cell_cta")  # small grids: four warps share a cell (barriers between the phases)
    assert _count(cta, "VABSDIFF4") >= 4 and _count(cta, "VIMNMX3") >= 40 and _count(cta, "BAR") >= 4
    assert _count(cta, "ACQBULK") == 1 and _count(cta, "LDL") == 0 and _count(cta, "STL") == 0
    # the cell / level tables are constants and may be read before the wait; the window rows (128-bit loads) may not
    first_ld = min(i for i, o in enumerate(cta) if o.startswith("LDG") and ".128" in o)
    assert cta.index("ACQBULK") < first_ld
    tma = _kernel(kernels, "k_fast_cellsILb0ELb1E")
    assert _count(tma, "UTMALDG") == 1  # the opt-in TMA staging variant
