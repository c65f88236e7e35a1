// Warp-level exact re-rank primitives shared by the neighbour-search kernels:
// (squared distance, sample index) pairs, one or more per lane, ordered lexicographically.
#pragma once

#include "common.cuh"

namespace cmx {

__device__ __forceinline__ bool lex_less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}

// ascending bitonic sort of one (d, i) element per lane
__device__ __forceinline__ void warp_sort32(double& d, int& i, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const double pd = __shfl_xor_sync(kFull, d, j);
            const int pi = __shfl_xor_sync(kFull, i, j);
            const bool up = (lane & k2) == 0;
            const bool lower = (lane & j) == 0;
            const bool p_less = lex_less(pd, pi, d, i);
            const bool take = (lower == up) ? p_less : !p_less;
            if (take) {
                d = pd;
                i = pi;
            }
        }
    }
}

// list: KP = 32*E elements, element e*32+lane in (ld[e], li[e]), ascending.
// (cd, ci): 32 sorted candidates, one per lane.  Keeps the KP smallest, sorted.
template <int E>
__device__ __forceinline__ void merge_sorted32(double (&ld)[E], int (&li)[E], double cd, int ci, int lane) {
    // min(list[i], cand_reversed[i]) over the top 32 slots -> bitonic sequence of the KP smallest
    const double rd = __shfl_sync(kFull, cd, 31 - lane);
    const int ri = __shfl_sync(kFull, ci, 31 - lane);
    if (lex_less(rd, ri, ld[E - 1], li[E - 1])) {
        ld[E - 1] = rd;
        li[E - 1] = ri;
    }
#pragma unroll
    for (int j = E / 2; j >= 1; j >>= 1) {  // strides of j*32 elements stay inside a lane
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if ((e & j) == 0 && e + j < E) {
                if (lex_less(ld[e + j], li[e + j], ld[e], li[e])) {
                    const double td = ld[e];
                    const int ti = li[e];
                    ld[e] = ld[e + j];
                    li[e] = li[e + j];
                    ld[e + j] = td;
                    li[e + j] = ti;
                }
            }
        }
    }
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double pd = __shfl_xor_sync(kFull, ld[e], j);
            const int pi = __shfl_xor_sync(kFull, li[e], j);
            const bool lower = (lane & j) == 0;
            const bool p_less = lex_less(pd, pi, ld[e], li[e]);
            if (lower ? p_less : !p_less) {
                ld[e] = pd;
                li[e] = pi;
            }
        }
    }
}


}  // namespace cmx
