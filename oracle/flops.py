"""ORACLE / measurement helper -- test and bench infrastructure only.

Algorithmic conv FLOPs per event clip (2*MAC over every Conv3d on the forward_features
path at the reference's un-padded shapes), the figure SURVEY.md section 8d fixes for
`roofline.achieved` (100.615 GFLOP for SlowFast-R50 8x8 at 224x224)."""
from __future__ import annotations


def _out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def conv_gflop_per_clip(spec, crop: int = 224) -> float:
    total = 0.0
    dims = []
    for p, t in enumerate(spec.pathway_frames()):
        st = spec.stems[p]
        c = st.conv
        t1, h1, w1 = (_out(t, c.kernel[0], c.stride[0], c.pad[0]), _out(crop, c.kernel[1], c.stride[1], c.pad[1]),
                      _out(crop, c.kernel[2], c.stride[2], c.pad[2]))
        total += t1 * h1 * w1 * c.flops_per_out_pixel
        dims.append([_out(t1, st.pool_kernel[0], st.pool_stride[0], st.pool_pad[0]),
                     _out(h1, st.pool_kernel[1], st.pool_stride[1], st.pool_pad[1]),
                     _out(w1, st.pool_kernel[2], st.pool_stride[2], st.pool_pad[2])])
    for si in range(4):
        f = spec.fuses[si]
        if f is not None:
            t, h, w = dims[1]
            total += _out(t, f.kernel[0], f.stride[0], f.pad[0]) * h * w * f.flops_per_out_pixel
        for p, blocks in enumerate(spec.stages[si]):
            for b in blocks:
                t, h, w = dims[p]
                for c in (b.a, b.b):
                    t, h, w = (_out(t, c.kernel[0], c.stride[0], c.pad[0]), _out(h, c.kernel[1], c.stride[1], c.pad[1]),
                               _out(w, c.kernel[2], c.stride[2], c.pad[2]))
                    total += t * h * w * c.flops_per_out_pixel
                total += t * h * w * b.c.flops_per_out_pixel
                if b.branch1 is not None:
                    total += t * h * w * b.branch1.flops_per_out_pixel
                dims[p] = [t, h, w]
                if b.nonlocal_ is not None:
                    nl = b.nonlocal_
                    tp, hp, wp = (t, h, w) if nl.pool is None else (t // nl.pool[0], h // nl.pool[1], w // nl.pool[2])
                    total += t * h * w * (nl.theta.flops_per_out_pixel + nl.out.flops_per_out_pixel)
                    total += tp * hp * wp * (nl.phi.flops_per_out_pixel + nl.g.flops_per_out_pixel)
        if si == 0:
            for p in range(len(dims)):
                k = spec.pool1[p]
                dims[p] = [dims[p][0] // k[0], dims[p][1] // k[1], dims[p][2] // k[2]]
    return total / 1e9
