// orbb_wire.cu -- wire format of a processed frame: the BSON document WebSocketCom sends to the UI
// (reference src/WebSocket/WebSocketCom.cpp:164-184, writer src/WebSocket/bson.cpp:46-130).  Host-only byte packing:
//   int32 total size | elements | 0x00
//   element = type byte | key, 0-terminated | value;  int32 value = 4 bytes LE;  binary value = int32 length |
//   subtype 0x80 | bytes.  Keys in the reference's order: ax ay az width height channels keypoints_x keypoints_y image.
// Written straight into the caller's buffer in one pass (the reference collects items in a std::vector, mallocs the
// message and copies every payload twice).  Checked byte for byte against the reference's own Bson class
// (oracle/_ref/libref_bson.so, tests/test_wire_format.py).
#include <cstring>

#include "../../include/orbb200.h"

namespace {
struct Writer {
    uint8_t *p;
    void i32(const char *key, int32_t v) {
        *p++ = 0x10;
        const size_t k = std::strlen(key) + 1;
        std::memcpy(p, key, k); p += k;
        std::memcpy(p, &v, 4); p += 4;
    }
    void bin(const char *key, const void *data, uint32_t n) {
        *p++ = 0x05;
        const size_t k = std::strlen(key) + 1;
        std::memcpy(p, key, k); p += k;
        std::memcpy(p, &n, 4); p += 4;
        *p++ = 0x80;
        if (n) std::memcpy(p, data, n);
        p += n;
    }
};
// type byte + key + terminator
constexpr size_t kKeys = (1 + 3) * 3 + (1 + 6) + (1 + 7) + (1 + 9) + (1 + 12) * 2 + (1 + 6);
}  // namespace

extern "C" size_t orbb_slam_frame_bson_size(int n_matched, size_t image_bytes) {
    if (n_matched < 0) return 0;
    return 4 + kKeys + 6 * 4 + 3 * (4 + 1) + 2 * (size_t)n_matched * sizeof(uint16_t) + image_bytes + 1;
}

extern "C" long long orbb_slam_frame_to_bson(int32_t ax, int32_t ay, int32_t az, int32_t width, int32_t height,
                                             int32_t channels, const uint16_t *keypoints_x, const uint16_t *keypoints_y,
                                             int n_matched, const uint8_t *image, size_t image_bytes, uint8_t *out,
                                             size_t out_capacity) {
    if (!out || n_matched < 0 || (n_matched > 0 && (!keypoints_x || !keypoints_y)) || (image_bytes > 0 && !image))
        return ORBB_ERR_INVALID;
    const size_t total = orbb_slam_frame_bson_size(n_matched, image_bytes);
    if (total > 0xffffffffull) return ORBB_ERR_CAPACITY;
    if (total > out_capacity) return ORBB_ERR_CAPACITY;
    Writer w{out};
    const uint32_t t32 = (uint32_t)total;
    std::memcpy(w.p, &t32, 4); w.p += 4;
    w.i32("ax", ax); w.i32("ay", ay); w.i32("az", az);
    w.i32("width", width); w.i32("height", height); w.i32("channels", channels);
    w.bin("keypoints_x", keypoints_x, (uint32_t)(n_matched * sizeof(uint16_t)));
    w.bin("keypoints_y", keypoints_y, (uint32_t)(n_matched * sizeof(uint16_t)));
    w.bin("image", image, (uint32_t)image_bytes);
    *w.p++ = 0;
    return (long long)(w.p - out);
}
