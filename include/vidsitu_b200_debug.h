/*
 * Debug-only entry points of libvidsitu_b200.so (not part of the drop-in
 * boundary; used by tools/gpu_probe.py to pin hardware semantics).
 */
#ifndef VIDSITU_B200_DEBUG_H_
#define VIDSITU_B200_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif

/* One im2col-mode TMA load (SWIZZLE_NONE) of a bf16 [n,t,h,w,pitch] tensor with the
 * given corners / traversal strides / box, issued at base coordinate
 * (cc,cw,ch,cd,cn) with filter offsets (ow,oh,od).  `out` receives the
 * pixel_box * chan_box bf16 values exactly as they landed in shared memory. */
int vsb_debug_im2col_probe(const void* in, int n, int t, int h, int w, int c, int pitch, int lw, int lh, int lt,
                           int uw, int uh, int ut, int sw, int sh, int st, int chan_box, int pixel_box, int cc,
                           int cw, int ch, int cd, int cn, int ow, int oh, int od, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
