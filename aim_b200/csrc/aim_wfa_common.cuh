// Device helpers shared by the WFA kernels (aim_wfa.cu: one pair per warp; aim_wfa_sub.cu: several
// pairs per warp in lockstep).
#ifndef AIM_WFA_COMMON_CUH
#define AIM_WFA_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace aim {
namespace {

constexpr int kNull = -16384;  // AFFINE_WAVEFRONT_OFFSET_NULL = INT16_MIN / 2 (WFA/DPU-MRAM/common/common.h:95)
constexpr unsigned kFull = 0xffffffffu;

// 8 ASCII bases -> 16 bits, first base in the two most significant bits.  (c >> 1) & 3 maps
// A,C,G,T to 0,1,3,2.  *ok is cleared if any of the first `valid` bytes is not one of A,C,G,T.
__device__ __forceinline__ uint32_t pack8(uint2 w, int valid, bool *ok)
{
    uint32_t out = 0;
    uint32_t words[2] = {w.x, w.y};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t x = words[j];
        int nv = valid - 4 * j;  // bytes of this word that belong to the sequence
        uint32_t keep = nv >= 4 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u));
        x = (x & keep) | (0x41414141u & ~keep);  // pad with 'A'
        uint32_t c = (x >> 1) & 0x03030303u;
        uint32_t is2 = (c >> 1) & ~c & 0x01010101u;  // code 2 <=> 'T' (0x54 = 0x41 + 2*2 + 15)
        uint32_t expect = 0x41414141u + 2u * c + 15u * is2;
        if (expect != x) *ok = false;
        uint32_t b = ((c << 6) | (c >> 4) | (c >> 14) | (c >> 24)) & 0xffu;
        out = (out << 8) | b;
    }
    return out;
}

// Number of equal bases from pattern[v], text[h], at most lim (> 0); 2-bit packed shared memory.
__device__ __forceinline__ int match_packed(const uint32_t *sP, const uint32_t *sT, int v, int h, int lim)
{
    int cnt = 0;
    for (;;) {
        int pv = v + cnt, ph = h + cnt;
        uint32_t a = __funnelshift_l(sP[(pv >> 4) + 1], sP[pv >> 4], (pv & 15) * 2);
        uint32_t b = __funnelshift_l(sT[(ph >> 4) + 1], sT[ph >> 4], (ph & 15) * 2);
        uint32_t d = a ^ b;
        if (d) { cnt += __clz(d) >> 1; break; }
        cnt += 16;
        if (cnt >= lim) break;
    }
    return min(cnt, lim);
}
__device__ __forceinline__ int match_bytes(const char *gp, const char *gt, int v, int h, int lim)
{
    int cnt = 0;
    while (cnt < lim && gp[v + cnt] == gt[h + cnt]) ++cnt;
    return cnt;
}

__device__ __forceinline__ int warp_min(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

}  // namespace
}  // namespace aim
#endif
