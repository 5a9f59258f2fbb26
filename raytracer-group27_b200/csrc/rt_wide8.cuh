// Eight lanes per ray through the 8-wide tree (see rt_wide8.cu for the why and the node layout): the per-group traversal.
#pragma once
#include "rt_trace.cuh"

namespace rtb {

#ifndef RT_WIDE_NEAREST_ONLY
#define RT_WIDE_NEAREST_ONLY 1
#endif
constexpr int kGroup = 8;                 // lanes per ray
constexpr int kWideBlock = 128;           // 16 groups per block
constexpr int kWideStack = 64;            // entries per group (shared memory); a step pushes at most 7, the tree is at most 8 levels deep for 2^22 triangles

// One group of eight lanes walks one query through the 8-wide tree.  All arguments and the result are uniform across the group.
template <bool ANYHIT>
__device__ __forceinline__ void trace_wide(const SceneDev& s, const float4* __restrict__ wide, int root_entry, const f3& o, const f3& d, HitRec& best, int* stack)
{
    const int lane = threadIdx.x & 31, sub = lane & (kGroup - 1);
    const unsigned gbase = (unsigned)(lane & ~(kGroup - 1)), gmask = 0xffu << gbase;
    const f3 dn = xnormalize(d);
    const float rx = fabsf(dn.x) > 1e-30f ? 1.0f / dn.x : copysignf(1e30f, dn.x);
    const float ry = fabsf(dn.y) > 1e-30f ? 1.0f / dn.y : copysignf(1e30f, dn.y);
    const float rz = fabsf(dn.z) > 1e-30f ? 1.0f / dn.z : copysignf(1e30f, dn.z);
    const float nlx = rx >= 0.0f ? rx : 0.0f, nhx = rx >= 0.0f ? 0.0f : rx;
    const float nly = ry >= 0.0f ? ry : 0.0f, nhy = ry >= 0.0f ? 0.0f : ry;
    const float nlz = rz >= 0.0f ? rz : 0.0f, nhz = rz >= 0.0f ? 0.0f : rz;
    const float oix = o.x * rx, oiy = o.y * ry, oiz = o.z * rz;
    float tlimit = prune_limit(best.t);
    TraceStats st;
    int cur = root_entry, sp = 0;
    if (s.n_spheres > 0 && test_spheres(s, o, d, best)) {
        tlimit = prune_limit(best.t);
        if (ANYHIT)
            cur = kTravDone;
    }
    while (cur != kTravDone) {
        if (cur >= 0) { // node: every lane tests one child
            const float4* np = wide + 16 * (size_t)RT_GUARD(cur, s.n_wide_nodes, kChkWideNode) + 2 * sub;
            float4 a0, a1; // this lane's child: one 32-byte load
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w), "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w)
                         : "l"(np));
            const float nx = fmaf(a0.x, nlx, fmaf(a1.x, nhx, -oix)), fx = fmaf(a0.x, nhx, fmaf(a1.x, nlx, -oix));
            const float ny = fmaf(a0.y, nly, fmaf(a1.y, nhy, -oiy)), fy = fmaf(a0.y, nhy, fmaf(a1.y, nly, -oiy));
            const float nz = fmaf(a0.z, nlz, fmaf(a1.z, nhz, -oiz)), fz = fmaf(a0.z, nhz, fmaf(a1.z, nlz, -oiz));
            const float tn = fmaxf(fmaxf(fmaxf(nx, ny), nz), 0.0f), tf = fminf(fminf(fminf(fx, fy), fz), tlimit);
            const bool hit = tn <= tf * 1.0000005f; // an inside-out box (missing child) gives tn = +big, tf = -big: no hit
            const unsigned hm = (__ballot_sync(gmask, hit) >> gbase) & 0xffu;
            if (hm == 0u) {
                cur = sp ? stack[--sp] : kTravDone;
                continue;
            }
            const int n = __popc(hm);
            int rank; // position of this child among the hit ones: nearest first (ties: lower child index); any order will do for an any-hit query
            if (ANYHIT) {
                rank = __popc(hm & ((1u << sub) - 1u));
#if RT_WIDE_NEAREST_ONLY
            } else if (true) { // only the nearest child is singled out (one REDUX on (entry distance, child index) keys), the others keep child order:
                               // 18 % faster than ranking all hit children through eight shuffles (tools/wide8_probe.py: 16 K rays 81 -> 67 us, 1 M 1640 -> 1344 us)
                const unsigned key = hit ? ((__float_as_uint(tn) & ~7u) | (unsigned)sub) : 0xffffffffu;
                const unsigned kmin = __reduce_min_sync(gmask, key);
                const int nearest = (int)(kmin & 7u);
                const int below = __popc(hm & ((1u << sub) - 1u)); // hit children with a lower index
                rank = sub == nearest ? 0 : below + (sub < nearest ? 1 : 0);
#endif
            } else {
                rank = 0;
#pragma unroll
                for (int j = 0; j < kGroup; j++) {
                    const float tj = __shfl_sync(gmask, tn, (int)gbase + j);
                    rank += (((hm >> j) & 1u) && (tj < tn || (tj == tn && j < sub))) ? 1 : 0;
                }
            }
            const int entry = __float_as_int(a0.w);
            const unsigned first = (__ballot_sync(gmask, hit && rank == 0) >> gbase) & 0xffu;
            const int next = __shfl_sync(gmask, entry, (int)gbase + __ffs(first) - 1);
            if (hit && rank > 0)
                stack[RT_GUARD(sp + (n - 1 - rank), kWideStack, kChkWideStack)] = entry; // the farthest lowest, the second nearest on top
            sp += n - 1;
            __syncwarp(gmask);
            cur = next;
        } else { // leaf: every lane tests one triangle; the group keeps the lexicographic minimum of (t, tie key)
            const int enc = ~cur;
            const int first = enc >> 3, count = (enc & 7) + 1;
            HitRec mine = best;
            bool found = false;
            if (sub < count)
                found = test_triangle<false>(s, first + sub, o, d, dn, mine, st);
            if (!found)
                mine = HitRec { FLT_MAX, INT_MAX, -1 };
            const unsigned any = __ballot_sync(gmask, found);
            if (any) {
#pragma unroll
                for (int off = kGroup / 2; off > 0; off >>= 1) {
                    const float t2 = __shfl_xor_sync(gmask, mine.t, off);
                    const int k2 = __shfl_xor_sync(gmask, mine.key, off), i2 = __shfl_xor_sync(gmask, mine.ti, off);
                    if (t2 < mine.t || (t2 == mine.t && k2 < mine.key))
                        mine = HitRec { t2, k2, i2 };
                }
                best = mine;
                tlimit = prune_limit(best.t);
                if (ANYHIT) {
                    cur = kTravDone;
                    continue;
                }
            }
            cur = sp ? stack[--sp] : kTravDone;
        }
    }
}

} // namespace rtb
