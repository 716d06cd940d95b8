// C++ host-mirror check: runs orbb200::ORBextractor::operator() on a raw gray frame read from argv[1]
// (w h on the command line) and writes keypoints + descriptors as raw bytes to argv[4], so the pytest
// GPU tier can compare them with the ctypes path / the oracle.  Also exercises the Jetracer:: stage names.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orbb200.hpp"

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s frame.raw w h out.bin\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]);
    std::vector<uint8_t> img((size_t)w * h);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "bad input\n"); return 2; }
    std::fclose(f);
    try {
        orbb200::ORBextractor ex(1000, 1.2f, 8, 20, 7, w, h);
        std::vector<orbb200::KeyPoint> kp;
        std::vector<uint8_t> desc;
        ex(img.data(), (size_t)w, nullptr, kp, desc);
        if (ex.GetLevels() != 8 || ex.GetScaleFactors().size() != 8) return 3;
        FILE *o = std::fopen(argv[4], "wb");
        const int n = (int)kp.size();
        std::fwrite(&n, sizeof(int), 1, o);
        std::fwrite(kp.data(), sizeof(orbb200::KeyPoint), kp.size(), o);
        std::fwrite(desc.data(), 1, desc.size(), o);
        std::fclose(o);
        std::printf("%d keypoints\n", n);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
