// Warp-level streaming top-k ("smallest k of a stream") held in registers.
//
// One warp owns one query. The running result is a list of 32*KPL (distance, index) pairs,
// sorted ascending over rank r = slot*32 + lane and kept entirely in registers. Candidates that
// beat the current k-th entry are compacted into a small shared-memory queue; every 32 accepted
// candidates are bitonic-sorted across the lanes with shuffles and bitonic-merged into the list.
// Ordering is the total order (distance, index): ties in distance go to the lower index, so the
// result does not depend on scheduling. NaN distances never enter the list.
#pragma once
#include "fs_common.cuh"

__device__ __forceinline__ bool fs_pair_less(float ad, int ai, float bd, int bi) {
    return ad < bd || (ad == bd && ai < bi);
}

// Ascending bitonic sort of one (d, i) pair per lane.
__device__ __forceinline__ void fs_warp_bitonic_sort(float& d, int& i, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            float od = __shfl_xor_sync(FS_FULL_MASK, d, j);
            int oi = __shfl_xor_sync(FS_FULL_MASK, i, j);
            bool up = (lane & k) == 0 || k == 32;
            bool lower = (lane & j) == 0;
            bool keep_min = (lower == up);
            bool o_less = fs_pair_less(od, oi, d, i);
            bool take = keep_min ? o_less : fs_pair_less(d, i, od, oi);
            if (take) { d = od; i = oi; }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Packed 64-bit keys: (order-preserving float bits << 32) | index. With unique indices the keys are unique, so a
// compare-exchange is one unsigned 64-bit comparison instead of the two-level (distance, index) test.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long fs_pack_key(float d, int i) {
    const uint32_t b = __float_as_uint(d);
    const uint32_t o = b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u);
    return ((unsigned long long)o << 32) | (uint32_t)i;
}
__device__ __forceinline__ void fs_unpack_key(unsigned long long key, float& d, int& i) {
    const uint32_t o = (uint32_t)(key >> 32);
    d = __uint_as_float(o ^ ((o & 0x80000000u) ? 0x80000000u : 0xffffffffu));
    i = (int)(uint32_t)key;
}
// Ascending bitonic sort of 32*H keys, element e = h*32 + lane (H = 1, 2 or 4).
template <int H>
__device__ __forceinline__ void fs_warp_bitonic_sort_keys(unsigned long long (&key)[H], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * H; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {                       // partner element lives in the same lane: slot h ^ (j / 32)
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const int hp = h | (j >> 5);
                    if ((h & (j >> 5)) == 0 && hp < H) {
                        const bool up = ((h * 32) & k) == 0 || k == 32 * H;      // lane bits do not reach k here
                        const bool swap = up ? key[hp] < key[h] : key[h] < key[hp];
                        if (swap) { const unsigned long long t = key[h]; key[h] = key[hp]; key[hp] = t; }
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const unsigned long long o = __shfl_xor_sync(FS_FULL_MASK, key[h], j);
                    const int e = h * 32 + lane;
                    const bool up = (e & k) == 0 || k == 32 * H;
                    const bool lower = (lane & j) == 0;
                    const bool keep_min = (lower == up);
                    if ((o < key[h]) == keep_min) key[h] = o;
                }
            }
        }
    }
}

template <int KPL>
struct FsWarpSelect {
    static_assert(KPL == 1 || KPL == 2 || KPL == 4, "list length must be 32, 64 or 128");
    float d[KPL];
    int i[KPL];
    float thr_d;
    int thr_i;
    int qn;          // queue fill, warp-uniform
    float* qd;       // shared-memory queue, 64 entries per warp
    int* qi;
    int kk;          // number of entries wanted (<= 32*KPL)
    int lane;

    __device__ __forceinline__ void init(float* queue_d, int* queue_i, int want) {
        lane = threadIdx.x & 31;
        qd = queue_d;
        qi = queue_i;
        kk = want;
        qn = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) { d[s] = INFINITY; i[s] = FS_IDX_PAD; }
        thr_d = INFINITY;
        thr_i = FS_IDX_PAD;
    }

    __device__ __forceinline__ void refresh_threshold() {
        const int r = kk - 1;
        float td = d[0];
        int ti = i[0];
#pragma unroll
        for (int s = 1; s < KPL; ++s)
            if ((r >> 5) == s) { td = d[s]; ti = i[s]; }
        thr_d = __shfl_sync(FS_FULL_MASK, td, r & 31);
        thr_i = __shfl_sync(FS_FULL_MASK, ti, r & 31);
    }

    // Merge the first min(qn, 32) queue entries into the sorted list.
    __device__ __forceinline__ void flush() {
        const int take = qn < 32 ? qn : 32;
        float nd = INFINITY;
        int ni = FS_IDX_PAD;
        if (lane < take) { nd = qd[lane]; ni = qi[lane]; }
        float rd = INFINITY;
        int ri = FS_IDX_PAD;
        if (qn > 32 && lane + 32 < qn) { rd = qd[lane + 32]; ri = qi[lane + 32]; }
        __syncwarp();
        if (qn > 32 && lane + 32 < qn) { qd[lane] = rd; qi[lane] = ri; }
        qn = qn > 32 ? qn - 32 : 0;
        __syncwarp();

        fs_warp_bitonic_sort(nd, ni, lane);
        // descending copy of the new batch against the last (largest) slot of the ascending list
        float xd = __shfl_sync(FS_FULL_MASK, nd, 31 - lane);
        int xi = __shfl_sync(FS_FULL_MASK, ni, 31 - lane);
        if (fs_pair_less(xd, xi, d[KPL - 1], i[KPL - 1])) { d[KPL - 1] = xd; i[KPL - 1] = xi; }
        // the list is now bitonic over 32*KPL ranks: bitonic merge
#pragma unroll
        for (int st = KPL / 2; st > 0; st >>= 1) {
#pragma unroll
            for (int s = 0; s < KPL; ++s) {
                if ((s & st) == 0) {
                    if (fs_pair_less(d[s + st], i[s + st], d[s], i[s])) {
                        float td = d[s]; d[s] = d[s + st]; d[s + st] = td;
                        int ti = i[s]; i[s] = i[s + st]; i[s + st] = ti;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
#pragma unroll
            for (int s = 0; s < KPL; ++s) {
                float od = __shfl_xor_sync(FS_FULL_MASK, d[s], j);
                int oi = __shfl_xor_sync(FS_FULL_MASK, i[s], j);
                bool lower = (lane & j) == 0;
                bool o_less = fs_pair_less(od, oi, d[s], i[s]);
                bool take_o = lower ? o_less : fs_pair_less(d[s], i[s], od, oi);
                if (take_o) { d[s] = od; i[s] = oi; }
            }
        }
        refresh_threshold();
    }

    // Offer one candidate per lane (all 32 lanes must call; `valid` masks tail lanes).
    __device__ __forceinline__ void offer(float cd, int ci, bool valid) {
        bool acc = valid && fs_pair_less(cd, ci, thr_d, thr_i);
        unsigned m = __ballot_sync(FS_FULL_MASK, acc);
        if (m) {
            int pos = qn + __popc(m & ((1u << lane) - 1u));
            if (acc) { qd[pos] = cd; qi[pos] = ci; }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) flush();
        }
    }

    __device__ __forceinline__ void finish() {
        if (qn > 0) flush();
    }

    // rank r -> (distance, index); valid on every lane after the call.
    __device__ __forceinline__ void get(int r, float& od, int& oi) const {
        float td = d[0];
        int ti = i[0];
#pragma unroll
        for (int s = 1; s < KPL; ++s)
            if ((r >> 5) == s) { td = d[s]; ti = i[s]; }
        od = __shfl_sync(FS_FULL_MASK, td, r & 31);
        oi = __shfl_sync(FS_FULL_MASK, ti, r & 31);
    }

    // Coalesced write of ranks [skip, kk) to out_i / out_d (nullable), remapping pad entries.
    __device__ __forceinline__ void store(int skip, int* out_i, float* out_d, int idx_offset,
                                          int pad_index, float pad_dist) const {
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            int r = s * 32 + lane;
            if (r >= skip && r < kk) {
                int ii = i[s];
                float dd = d[s];
                if (ii == FS_IDX_PAD) { ii = pad_index; dd = pad_dist; } else { ii += idx_offset; }
                out_i[r - skip] = ii;
                if (out_d) out_d[r - skip] = dd;
            }
        }
    }
};
