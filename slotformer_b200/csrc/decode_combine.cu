// Decoder epilogue for sm_100a: everything after the deconvolution stack of
// the spatial-broadcast decoder, in one streaming pass.
//
// Replaces the tail of reference StoSAVi.decode (base_slots/models/savi.py:519-523)
//     recons, masks = out[:, :, :3], softmax(out[:, :, 3:], dim=1);  recon_combined = sum_k recons * masks
// and reference postproc_mask (video_prediction/vp_utils.py:20-41: per-frame background slot = the slot whose
// largest mask value is smallest; pixels whose best score is below FG_THRE go to it; argmax over slots).
//
// HBM-bound, no reuse: decode_combine_kernel reads the K x 4 planes of a frame once (float4 = 4 pixels per
// thread, every load of a thread issued before the first use) and writes K mask planes + 3 colour planes;
// algorithmic bytes per pixel = 16 K (read) + 4 K + 12 (write).  The per-(frame, slot) mask maxima needed by
// postproc are reduced on the fly (warp shuffle + one atomicMax per warp on the non-negative float's bit
// pattern), so seg_argmax_kernel reads only the K mask planes again.
#include "common.cuh"
#include "decode_kernel.h"

namespace sfb {

// Order-preserving float <-> uint32 map (any sign): u(a) < u(b) <=> a < b, and 0 is below every encoded value, so a
// zero-filled word is the identity of atomicMax.  postproc_mask accepts arbitrary scores, not only probabilities.
__device__ __forceinline__ unsigned int ord_enc(float v) {
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}


static constexpr int DC_THREADS = 256;
static constexpr int DC_MAXK = 16;

template <int K>
__global__ void __launch_bounds__(DC_THREADS) decode_combine_kernel(const float* __restrict__ dec, float* __restrict__ masks,
                                                                    float* __restrict__ recon, unsigned int* __restrict__ slot_max,
                                                                    int HW4) {
    const int b = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(dec) + (size_t)b * K * 4 * HW4;
    float4* mk = reinterpret_cast<float4*>(masks) + (size_t)b * K * HW4;
    float4* rc = reinterpret_cast<float4*>(recon) + (size_t)b * 3 * HW4;
    float vmax[K];
#pragma unroll
    for (int k = 0; k < K; ++k) vmax[k] = 0.f;
    for (int i = blockIdx.x * DC_THREADS + threadIdx.x; i < HW4; i += gridDim.x * DC_THREADS) {
        float4 lg[K], r[K], g[K], bl[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4* pk = src + (size_t)k * 4 * HW4 + i;
            r[k] = __ldcs(pk); g[k] = __ldcs(pk + HW4); bl[k] = __ldcs(pk + 2 * HW4); lg[k] = __ldcs(pk + 3 * HW4);
        }
        float4 m = lg[0];
#pragma unroll
        for (int k = 1; k < K; ++k) {
            m.x = fmaxf(m.x, lg[k].x); m.y = fmaxf(m.y, lg[k].y); m.z = fmaxf(m.z, lg[k].z); m.w = fmaxf(m.w, lg[k].w);
        }
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            lg[k].x = expf(lg[k].x - m.x); lg[k].y = expf(lg[k].y - m.y);
            lg[k].z = expf(lg[k].z - m.z); lg[k].w = expf(lg[k].w - m.w);
            s.x += lg[k].x; s.y += lg[k].y; s.z += lg[k].z; s.w += lg[k].w;
        }
        float4 cr = make_float4(0.f, 0.f, 0.f, 0.f), cg = cr, cb = cr;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float4 p;
            p.x = lg[k].x / s.x; p.y = lg[k].y / s.y; p.z = lg[k].z / s.z; p.w = lg[k].w / s.w;
            __stcs(mk + (size_t)k * HW4 + i, p);
            vmax[k] = fmaxf(vmax[k], fmaxf(fmaxf(p.x, p.y), fmaxf(p.z, p.w)));
            cr.x = fmaf(r[k].x, p.x, cr.x); cr.y = fmaf(r[k].y, p.y, cr.y); cr.z = fmaf(r[k].z, p.z, cr.z); cr.w = fmaf(r[k].w, p.w, cr.w);
            cg.x = fmaf(g[k].x, p.x, cg.x); cg.y = fmaf(g[k].y, p.y, cg.y); cg.z = fmaf(g[k].z, p.z, cg.z); cg.w = fmaf(g[k].w, p.w, cg.w);
            cb.x = fmaf(bl[k].x, p.x, cb.x); cb.y = fmaf(bl[k].y, p.y, cb.y); cb.z = fmaf(bl[k].z, p.z, cb.z); cb.w = fmaf(bl[k].w, p.w, cb.w);
        }
        __stcs(rc + i, cr); __stcs(rc + HW4 + i, cg); __stcs(rc + 2 * HW4 + i, cb);
    }
    if (slot_max != nullptr) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float v = vmax[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) atomicMax(slot_max + (size_t)b * K + k, ord_enc(v));
        }
    }
}

// max over the pixels of every (frame, slot) mask plane (stand-alone postproc_mask on given masks)
__global__ void __launch_bounds__(DC_THREADS) mask_max_kernel(const float* __restrict__ masks, unsigned int* __restrict__ slot_max,
                                                              int HW) {
    const float* mk = masks + (size_t)blockIdx.y * HW;
    float v = -INFINITY;
    for (int i = blockIdx.x * DC_THREADS + threadIdx.x; i < HW; i += gridDim.x * DC_THREADS) v = fmaxf(v, __ldg(mk + i));
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax(slot_max + blockIdx.y, ord_enc(v));
}

// one thread per pixel: first-occurrence argmin / argmax exactly as torch.argmin / torch.argmax define them
__global__ void __launch_bounds__(DC_THREADS) seg_argmax_kernel(const float* __restrict__ masks, const unsigned int* __restrict__ slot_max,
                                                                long long* __restrict__ seg, int K, int HW, float fg_thre) {
    const int b = blockIdx.y;
    __shared__ int bg_s;
    if (threadIdx.x == 0) {
        int bg = 0;
        float best = ord_dec(slot_max[(size_t)b * K]);
        for (int k = 1; k < K; ++k) {
            const float v = ord_dec(slot_max[(size_t)b * K + k]);
            if (v < best) { best = v; bg = k; }
        }
        bg_s = bg;
    }
    __syncthreads();
    const int bg = bg_s;
    const float* mk = masks + (size_t)b * K * HW;
    for (int i = blockIdx.x * DC_THREADS + threadIdx.x; i < HW; i += gridDim.x * DC_THREADS) {
        float v[DC_MAXK];
        float top = -INFINITY;
#pragma unroll
        for (int k = 0; k < DC_MAXK; ++k) {
            if (k < K) { v[k] = __ldg(mk + (size_t)k * HW + i); top = fmaxf(top, v[k]); }
        }
        const bool to_bg = top < fg_thre;
        int arg = 0;
        float best = -INFINITY;
#pragma unroll
        for (int k = 0; k < DC_MAXK; ++k) {
            if (k < K) {
                const float x = (to_bg && k == bg) ? 1.f : v[k];
                if (x > best) { best = x; arg = k; }
            }
        }
        seg[(size_t)b * HW + i] = arg;
    }
}

template <int K>
static cudaError_t combine_launch_t(const float* dec, float* masks, float* recon, unsigned int* slot_max, int B, int HW,
                                    int sms, cudaStream_t st) {
    const int HW4 = HW / 4;
    int bx = (HW4 + DC_THREADS - 1) / DC_THREADS;
    const int cap = (8 * sms + B - 1) / B;          // ~8 resident CTAs per SM over the whole launch
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    decode_combine_kernel<K><<<dim3(bx, B), DC_THREADS, 0, st>>>(dec, masks, recon, slot_max, HW4);
    return cudaGetLastError();
}

cudaError_t decode_combine_launch(const float* dec, float* masks, float* recon, unsigned int* slot_max, int B, int K, int HW,
                                  int sms, cudaStream_t st) {
    switch (K) {
        case 1: return combine_launch_t<1>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 2: return combine_launch_t<2>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 3: return combine_launch_t<3>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 4: return combine_launch_t<4>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 5: return combine_launch_t<5>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 6: return combine_launch_t<6>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 7: return combine_launch_t<7>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 8: return combine_launch_t<8>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 9: return combine_launch_t<9>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 10: return combine_launch_t<10>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 11: return combine_launch_t<11>(dec, masks, recon, slot_max, B, HW, sms, st);
        case 12: return combine_launch_t<12>(dec, masks, recon, slot_max, B, HW, sms, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t mask_max_launch(const float* masks, unsigned int* slot_max, int BK, int HW, int sms, cudaStream_t st) {
    int bx = (HW + 4 * DC_THREADS - 1) / (4 * DC_THREADS);
    const int cap = (8 * sms + BK - 1) / BK;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    mask_max_kernel<<<dim3(bx, BK), DC_THREADS, 0, st>>>(masks, slot_max, HW);
    return cudaGetLastError();
}

cudaError_t seg_argmax_launch(const float* masks, const unsigned int* slot_max, long long* seg, int B, int K, int HW,
                              float fg_thre, int sms, cudaStream_t st) {
    int bx = (HW + DC_THREADS - 1) / DC_THREADS;
    const int cap = (8 * sms + B - 1) / B;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    seg_argmax_kernel<<<dim3(bx, B), DC_THREADS, 0, st>>>(masks, slot_max, seg, K, HW, fg_thre);
    return cudaGetLastError();
}

}  // namespace sfb
