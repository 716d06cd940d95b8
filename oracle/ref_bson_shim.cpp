// ref_bson_shim.cpp -- REFERENCE-COMPILED ORACLE for the wire format.  TEST INFRASTRUCTURE ONLY.
//
// The one piece of the reference's frame path that compiles from its own sources without external libraries is its
// BSON writer (reference src/WebSocket/bson.{h,cpp}).  oracle/Makefile's `ref` target compiles THAT FILE where it lies
// under /root/reference together with this shim into oracle/_ref/libref_bson.so (git-ignored, never copied into the
// repo), so the product's packer (orbb_slam_frame_to_bson) is checked against bytes produced by the reference's own
// code.  The shim only (a) repeats the add() sequence of the reference's sender, WebSocketCom::handleEvent
// (src/WebSocket/WebSocketCom.cpp:164-184), and (b) supplies Bson::~Bson, which the reference defines inside
// WebSocketCom.cpp (:254-258, a file that needs websocketpp and does not build here).
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "bson.h"  // -I/root/reference/src/WebSocket

namespace Jetracer {
Bson::~Bson() {
    if (buffer_) free(buffer_);  // the reference pairs malloc with delete (WebSocketCom.cpp:256-257); same effect, defined
}
}  // namespace Jetracer

extern "C" size_t ref_slam_frame_bson(int32_t ax, int32_t ay, int32_t az, int32_t width, int32_t height,
                                      const uint16_t *keypoints_x, const uint16_t *keypoints_y, int32_t n_matched,
                                      const uint8_t *image, size_t image_length, uint8_t *out, size_t capacity) {
    using Jetracer::bson_value_type;
    Jetracer::Bson bson_message;
    int channels = 1;
    bson_message.add("ax", bson_value_type::bson_int32, &ax);
    bson_message.add("ay", bson_value_type::bson_int32, &ay);
    bson_message.add("az", bson_value_type::bson_int32, &az);
    bson_message.add("width", bson_value_type::bson_int32, &width);
    bson_message.add("height", bson_value_type::bson_int32, &height);
    bson_message.add("channels", bson_value_type::bson_int32, &channels);
    bson_message.add("keypoints_x", bson_value_type::bson_binary, const_cast<uint16_t *>(keypoints_x), n_matched * sizeof(uint16_t));
    bson_message.add("keypoints_y", bson_value_type::bson_binary, const_cast<uint16_t *>(keypoints_y), n_matched * sizeof(uint16_t));
    bson_message.add("image", bson_value_type::bson_binary, const_cast<uint8_t *>(image), image_length * sizeof(char));
    bson_message.process();
    const size_t n = bson_message.size();
    if (n <= capacity) std::memcpy(out, bson_message.ptr(), n);
    return n;
}
