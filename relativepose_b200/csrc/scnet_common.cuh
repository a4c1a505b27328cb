// scnet_common.cuh -- argument block shared by the float32 (scnet.cu) and tcgen05 (scnet_tc.cu) conv kernels.
#pragma once
#include <stdint.h>

#include "../../include/rp_b200.h"

namespace scnet {

constexpr float LEAKY = 0.1f;       // mymodel.py:20,33
constexpr int MAXTAP = 49;          // 7x7 stem of Resnet18_8s
constexpr double BN_EPS = 1e-5;

struct Tap { int dy, dx, widx; };
struct ConvClass { int py, px, Ha, Wb, ntap; Tap taps[MAXTAP]; };

struct ConvArgs {
    rp_conv_src src[2];
    int nsrc;
    int G, Hin, Win, Hout, Wout, Cout, Cin_total;
    int istr, ostr, nclass, tiles_m;
    int gsz;            // images per BN group
    ConvClass cls[4];
    const float* W;
    float* out; int out_pitch, out_ch_off;
    float* psum; float* psq;
    const float* bias; int tanh_out;
    int out_bf16;       // raw output stored as bfloat16 (tensor-core kernels, small-Cin stem)
};


extern long long g_conv_launches;

inline bool build_args(const rp_conv_desc* d, ConvArgs* A, int bm) {
    if (!d || d->nsrc < 1 || d->nsrc > 2 || d->k < 1 || d->k > 7 || d->s < 1 || d->s > 2 || d->G < 1) return false;
    if (d->transposed && d->k > 4) return false;
    A->gsz = d->imgs_per_group > 0 ? d->imgs_per_group : 2;
    A->nsrc = d->nsrc;
    A->Cin_total = 0;
    for (int i = 0; i < d->nsrc; ++i) { A->src[i] = d->src[i]; A->Cin_total += d->src[i].C; }
    A->G = d->G; A->Hin = d->Hin; A->Win = d->Win; A->Hout = d->Hout; A->Wout = d->Wout; A->Cout = d->Cout;
    A->W = d->W; A->out = d->out; A->out_pitch = d->out_pitch; A->out_ch_off = d->out_ch_off;
    A->psum = d->psum; A->psq = d->psq; A->bias = d->bias; A->tanh_out = d->tanh_out;
    A->out_bf16 = d->out_dtype == 1 ? 1 : 0;
    const int k = d->k, s = d->s, p = d->p;
    if (!d->transposed) {
        // iy = oy*s - p + ky
        A->istr = s; A->ostr = 1; A->nclass = 1;
        ConvClass& c = A->cls[0];
        c.py = 0; c.px = 0; c.Ha = d->Hout; c.Wb = d->Wout; c.ntap = 0;
        for (int ky = 0; ky < k; ++ky) for (int kx = 0; kx < k; ++kx) { c.taps[c.ntap].dy = ky - p; c.taps[c.ntap].dx = kx - p; c.taps[c.ntap].widx = ky * k + kx; ++c.ntap; }
    } else {
        // oy = iy*s - p + ky  ->  for output parity class py: ky with (py + p - ky) % s == 0, iy = a + (py + p - ky)/s
        A->istr = 1; A->ostr = s; A->nclass = s * s;
        for (int py = 0; py < s; ++py) for (int px = 0; px < s; ++px) {
            ConvClass& c = A->cls[py * s + px];
            c.py = py; c.px = px; c.ntap = 0;
            c.Ha = (d->Hout - py + s - 1) / s; c.Wb = (d->Wout - px + s - 1) / s;
            if (c.Ha < 0) c.Ha = 0; if (c.Wb < 0) c.Wb = 0;
            for (int ky = 0; ky < k; ++ky) {
                if (((py + p - ky) % s + s) % s != 0) continue;
                for (int kx = 0; kx < k; ++kx) {
                    if (((px + p - kx) % s + s) % s != 0) continue;
                    int dy = (py + p - ky) / s, dx = (px + p - kx) / s;     // exact (divisible); may be negative
                    if ((py + p - ky) < 0) dy = -((ky - py - p) / s);
                    if ((px + p - kx) < 0) dx = -((kx - px - p) / s);
                    c.taps[c.ntap].dy = dy; c.taps[c.ntap].dx = dx; c.taps[c.ntap].widx = ky * k + kx; ++c.ntap;
                }
            }
        }
    }
    int tm = 1;
    for (int i = 0; i < A->nclass; ++i) { int t = (A->gsz * A->cls[i].Ha * A->cls[i].Wb + bm - 1) / bm; if (t > tm) tm = t; }
    A->tiles_m = tm;
    return true;
}


}  // namespace scnet
