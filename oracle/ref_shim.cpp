// Test infrastructure only. extern "C" trampolines onto the reference's
// (C++-mangled) display-conversion functions declared in
// /root/reference/src/media/processing/yuvconversions.h:9-27, so that ctypes
// can call the UNMODIFIED reference object code. No arithmetic lives here.
#include "yuvconversions.h"
extern "C" {
int ref_yuv420_to_rgb_i_avx2_mt(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h, uint8_t t)
{ return yuv420_to_rgb_i_avx2_mt(in, out, w, h, t); }
int ref_yuv420_to_rgb_i_avx2(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h)
{ return yuv420_to_rgb_i_avx2(in, out, w, h); }
int ref_yuv420_to_rgb_i_sse41(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h)
{ return yuv420_to_rgb_i_sse41(in, out, w, h); }
void ref_yuv420_to_rgb_i_c(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h)
{ yuv420_to_rgb_i_c(in, out, w, h); }
void ref_half_rgb(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h)
{ half_rgb(in, out, w, h); }
void ref_flip_rgb(uint8_t* in, uint8_t* out, uint16_t w, uint16_t h, int hor, int ver)
{ flip_rgb(in, out, w, h, hor != 0, ver != 0); }
int ref_has_avx2(void) { return is_avx2_available() ? 1 : 0; }
}
