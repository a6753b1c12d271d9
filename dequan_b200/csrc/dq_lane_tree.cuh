// dq_lane_tree.cuh — lane-per-subtree COUNT_ALL search for small models whose pair filters are plain AND masks
// (at most 32 variables, no weak-equal / check-only entries: CompiledModel::small_ok && !has_f) — the generic path's
// answer to "one warp per subtree leaves most lanes idle": N-Queens with an extra constraint, small colouring or
// scheduling trees, anything the class engines do not recognise.
//
// One LANE owns one prefix subtree.  Its state, in shared memory as [word][thread] (conflict-free whatever a lane indexes):
//   R[l][q], q >= l   the domain of the variable at search position q once positions 0..l-1 are assigned — one row per
//                     level, so entering a level WRITES the next row (D &= mask, OpConstraint / AllDifferent::
//                     AplyArcConsistency -> Domain::Exclude / ExcludeInf / ExcludeSup, dequan.h:631-694, 915-939, 985-1172)
//                     and backtracking (RestoreSavedDomainStep, dequan.h:431-440) is just l - 1: no trail, no undo;
//   cand[l]           values of position l not tried yet | the value chosen << 27.
// The mask rows AND[x position][value][q] sit once per CTA in shared memory, transposed to [q][x][value] so that the
// lanes of a warp — all on the same q in a given trip of the row loop, each with its own (x, value) — spread over the banks.
// A trip of the loop tries ONE value per lane: a node (AssignVar, dequan.h:416-423) whether its check then passes or
// not; the row loop runs to the longest row among the warp's lanes.  Prefixes, partitions, first-solution keys and
// the result arrays are those of k_tree_dfs / k_tree_small (dq_kernels.cuh).
#pragma once
#include "dq_small_tree.cuh"

namespace dq {

constexpr int kLaneTreeThreads = 128;
constexpr int kLaneTreeRefill = 12;                  // idle lanes of a warp that trigger a refill
constexpr int kLaneTreeMaxDom = 27;                  // value index in 5 bits above a 27-bit untried mask

__host__ __device__ inline size_t lane_tree_smem(int nv, int kmax) {
    return ((size_t)nv * nv * kmax + ((size_t)nv * (nv + 1) / 2 + nv) * kLaneTreeThreads) * 4;
}

__global__ void __launch_bounds__(kLaneTreeThreads)
k_tree_lanes(SmallTablesDev T, TreeDfsArgs A) {
    extern __shared__ __align__(16) uint32_t lt_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nv = T.nv, kmax = T.kmax;
    uint32_t* s_tab = lt_raw;                                             // [q][x][value]
    uint32_t* rows = s_tab + (size_t)nv * nv * kmax;                      // [row entry][thread]
    uint32_t* cand = rows + (size_t)(nv * (nv + 1) / 2) * kLaneTreeThreads;   // [level][thread]
    for (int i = tid; i < nv * kmax * 32; i += kLaneTreeThreads) {
        const int q = i & 31, xb = i >> 5;                               // xb = x * kmax + value
        if (q < nv) s_tab[(size_t)q * nv * kmax + xb] = __ldg(T.t_and + i);
    }
    __syncthreads();
    auto R = [&](int l, int q) -> uint32_t& { return rows[(size_t)(l * nv - l * (l - 1) / 2 + q - l) * kLaneTreeThreads + tid]; };
    const unsigned long long slot = (unsigned long long)blockIdx.x * kLaneTreeThreads + tid;
    const uint32_t lt = (1u << lane) - 1u;
    const int d0 = A.depth;

    bool have = false, done = false, found = false;
    unsigned long long idx = 0, nodes = 0, sols = 0, acc_nodes = 0, acc_sols = 0;
    int l = 0;
    uint32_t c = 0;
    for (;;) {
        // ---- lanes without a subtree take the next prefixes of this partition ----
        // (in batches: replaying a prefix costs as much as a dozen nodes and idles the lanes that still have work)
        const uint32_t need = __ballot_sync(FULL, !have && !done);
        if (need && (__popc(need) >= kLaneTreeRefill || !__any_sync(FULL, have))) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(A.cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, 0);
            if (!have && !done) {
                idx = (base + __popc(need & lt)) * (unsigned long long)A.part_count + (unsigned long long)A.part_rank;
                if (idx >= A.n_prefix) done = true;
                else {
                    const uint8_t* prefix = A.prefixes + idx * (size_t)d0;
                    for (int q = 0; q < nv; q++) R(0, q) = __ldg(T.dom0_pos + q);
                    for (int i = 0; i < d0; i++) {
                        const int b = __ldg(prefix + i);
                        for (int q = i + 1; q < nv; q++) R(i + 1, q) = R(i, q) & s_tab[((size_t)q * nv + i) * kmax + b];
                    }
                    have = true; found = false; nodes = 0; sols = 0;
                    l = d0;
                    c = R(l, l);
                }
            }
            if (__all_sync(FULL, done)) break;
        }
        // ---- every value tried at this level: one level up, or the subtree is finished ----
        if (have && c == 0u) {
            if (l == d0) {
                A.sub_nodes[idx] = nodes;
                acc_nodes += nodes; acc_sols += sols;
                have = false;
            } else {
                --l;
                c = cand[(size_t)l * kLaneTreeThreads + tid] & 0x07FFFFFFu;
            }
        }
        // ---- AssignVar(next value) + forward check ----
        bool trying = have && c != 0u;
        const int b = __ffs((int)c) - 1;
        if (trying) {
            c &= c - 1u;
            ++nodes;
            if (l == nv - 1) {
                // the last variable: each remaining value is a node and, nothing being left to filter, a solution
                const uint32_t more = __popc(c);
                nodes += more;
                sols += 1u + more;
                if (!found) {
                    found = true;
                    const unsigned long long old = atomicMin(A.best_key, idx);
                    if (idx < old) {
                        uint8_t* out = A.sol + slot * nv;
                        const uint8_t* prefix = A.prefixes + idx * (size_t)d0;
                        for (int p = 0; p < nv; p++) {
                            const uint32_t v = p < d0 ? __ldg(prefix + p) : (p == nv - 1 ? (uint32_t)b : cand[(size_t)p * kLaneTreeThreads + tid] >> 27);
                            out[__ldg(T.order + p)] = (uint8_t)v;
                        }
                        A.sol_key[slot] = idx;
                    }
                }
                c = 0u;
                trying = false;
            }
        }
        const int n_rows = trying ? nv - 1 - l : 0;
        const int t_max = __reduce_max_sync(FULL, n_rows);
        bool wiped = false;
        const size_t xb = (size_t)l * kmax + (trying ? b : 0);
        for (int t = 0; t < t_max; t++) {
            const int q = nv - 1 - t;                                     // (the same q in every lane: see the table layout)
            if (t < n_rows) {
                const uint32_t nd = R(l, q) & s_tab[(size_t)q * nv * kmax + xb];
                R(l + 1, q) = nd;
                wiped |= nd == 0u;                                        // a wipe-out (dequan.h:663-668): this value fails
            }
        }
        if (trying && !wiped) {
            cand[(size_t)l * kLaneTreeThreads + tid] = c | ((uint32_t)b << 27);
            ++l;
            c = R(l, l);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc_nodes += __shfl_down_sync(FULL, acc_nodes, o);
        acc_sols += __shfl_down_sync(FULL, acc_sols, o);
    }
    if (lane == 0 && (acc_nodes | acc_sols)) {
        atomicAdd(A.totals + 0, acc_sols);
        atomicAdd(A.totals + 1, acc_nodes);
    }
}

}  // namespace dq
