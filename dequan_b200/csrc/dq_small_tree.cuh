// dq_small_tree.cuh — register-resident warp engine for single-tree solves of small models (<= 32 variables).
//
// The shape the north star describes, taken literally: one warp per prefix subtree, explicit-stack DFS, and the
// domain of the variable at search position p lives in a REGISTER of lane p (plus, for models with weak-equal or
// check-only constraints, its "fails validation" word F).  Assigning the variable at position d the value index b
// filters every unassigned variable at once:
//     D &= AND[d][b][lane]                       OpConstraint / AllDifferent::AplyArcConsistency -> Domain::Exclude /
//                                                 ExcludeInf / ExcludeSup (dequan.h:631-694, 915-939, 985-1172)
//     if (weq) D = (D & W) ? D & W : D, F = ~0   the weak Domain::Intersect of Equal (dequan.h:957-984)
//     F |= CHK[d][b][lane]                       check-only constraints (OrRange, user tables): Evaluate fails later
// with the tables in shared memory ([position][value][lane]: one conflict-free load per lane), the wipe-out test
// (dequan.h:663-668) one __any_sync, the value choice __ffs.  A failing value changes nothing; a passing one saves
// the lanes' registers in shared memory [depth][lane] — the per-depth saved-domain frame of
// Assignment::EnsureSavedDomain / RestoreSavedDomainStep (dequan.h:431-452) — so backtracking is one load per lane.
// Node accounting, FIRST/COUNT_ALL modes, prefix keys and partitioning are those of k_tree_dfs (dq_kernels.cuh).
#pragma once
#include "dq_kernels.cuh"

namespace dq {

struct SmallTablesDev {
    int nv, kmax;
    const uint32_t* __restrict__ t_and;      // [nv][kmax][32]
    const uint32_t* __restrict__ t_weq;      // [nv][kmax][32]   (HAS_F)
    const uint32_t* __restrict__ t_weq_on;   // [nv][kmax]       (HAS_F)
    const uint32_t* __restrict__ t_chk;      // [nv][kmax][32]   (HAS_F)
    const uint32_t* __restrict__ dom0_pos;   // [32] initial domain by position (all ones beyond nv)
    const uint16_t* __restrict__ order;      // [nv] position -> var id
};

__host__ __device__ inline size_t small_tree_smem(int nv, int kmax, bool has_f, int warps) {
    const size_t tab = (size_t)nv * kmax * 32 * 4;
    const size_t per_warp = (size_t)32 * 32 * 4 * (has_f ? 2 : 1) + 32 * 4 + 32;     // saved D (F), cand[32], val[32]
    return tab * (has_f ? 3 : 1) + (has_f ? (size_t)nv * kmax * 4 : 0) + per_warp * warps;   // AND (+ WEQ, CHK, WEQ-on) tables, per-warp frames
}

template <bool HAS_F>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_tree_small(SmallTablesDev T, TreeDfsArgs A) {
    extern __shared__ __align__(16) unsigned char st_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * kWarpsPerCta + wib;
    const int nv = T.nv, kmax = T.kmax;
    const size_t tab_words = (size_t)nv * kmax * 32;
    uint32_t* s_and = reinterpret_cast<uint32_t*>(st_raw);
    uint32_t* s_weq = s_and + tab_words;                         // HAS_F only
    uint32_t* s_chk = s_weq + (HAS_F ? tab_words : 0);
    uint32_t* s_weq_on = s_chk + (HAS_F ? tab_words : 0);
    uint32_t* per_warp = s_weq_on + (HAS_F ? (size_t)nv * kmax : 0);
    constexpr int kWarpWords = 32 * 32 * (HAS_F ? 2 : 1) + 32 + 8;
    uint32_t* saveD = per_warp + (size_t)wib * kWarpWords;       // [depth][lane]
    uint32_t* saveF = saveD + 32 * 32;                           // HAS_F only
    uint32_t* cand = saveD + 32 * 32 * (HAS_F ? 2 : 1);          // [depth] untried values
    uint8_t* val = reinterpret_cast<uint8_t*>(cand + 32);        // [depth] chosen value index

    for (size_t i = threadIdx.x; i < tab_words; i += blockDim.x) {
        s_and[i] = __ldg(T.t_and + i);
        if (HAS_F) { s_weq[i] = __ldg(T.t_weq + i); s_chk[i] = __ldg(T.t_chk + i); }
    }
    if (HAS_F) for (int i = threadIdx.x; i < nv * kmax; i += blockDim.x) s_weq_on[i] = __ldg(T.t_weq_on + i);
    __syncthreads();

    // forward check of "position d takes value index b" on this lane's variable; true = some domain is wiped out
    auto filter = [&](int d, int b, uint32_t D, uint32_t F, uint32_t& nd, uint32_t& nf) -> bool {
        const bool un = lane > d && lane < nv;                   // only unassigned variables are filtered (dequan.h:657-679)
        nd = D; nf = F;
        if (un) {
            const size_t at = ((size_t)d * kmax + b) * 32 + lane;
            nd = D & s_and[at];
            if (HAS_F) {
                if ((s_weq_on[d * kmax + b] >> lane) & 1u) {
                    const uint32_t m = s_weq[at];
                    if (nd & m) nd &= m; else nf = FULL;
                }
                nf |= s_chk[at];
            }
        }
        return __any_sync(FULL, un && nd == 0);
    };

    unsigned long long acc_nodes = 0, acc_sols = 0;
    for (;;) {
        unsigned long long j = 0;
        if (lane == 0) j = atomicAdd(A.cursor, 1ull);
        j = __shfl_sync(FULL, j, 0);
        const unsigned long long idx = j * (unsigned long long)A.part_count + (unsigned long long)A.part_rank;
        if (idx >= A.n_prefix) break;
        if (!A.count_all && *(volatile unsigned long long*)A.best_key < idx) break;   // later prefixes cannot win

        // ---- root state + the prefix's assignments ----
        uint32_t D = __ldg(T.dom0_pos + lane), F = 0;
        const uint8_t* prefix = A.prefixes + idx * (size_t)A.depth;
        for (int i = 0; i < A.depth; i++) {
            const int b = __ldg(prefix + i);
            if (lane == 0) val[i] = (uint8_t)b;
            uint32_t nd, nf;
            filter(i, b, D, F, nd, nf);
            D = nd; F = nf;
        }
        __syncwarp();

        // ---- explicit-stack DFS (ForwardCheckingStep, dequan.h:494-571) ----
        const int d0 = A.depth;
        unsigned long long nodes = 0, sols = 0;
        bool have_first = false;
        int outcome = 0;
        auto record_first = [&]() {
            unsigned long long old = 0;
            if (lane == 0) old = atomicMin(A.best_key, idx);
            old = __shfl_sync(FULL, old, 0);
            if (idx < old) {
                if (lane < nv) A.sol[(size_t)gw * nv + __ldg(T.order + lane)] = val[lane];
                if (lane == 0) A.sol_key[gw] = idx;
            }
        };
        if (d0 >= nv) { sols = 1; outcome = 1; have_first = true; }                    // IsComplete at entry, dequan.h:496-499
        else {
            int d = d0;
            uint32_t c = __shfl_sync(FULL, D, d);
            unsigned poll = 0;
            for (;;) {
                if (c == 0) {                                                          // every value tried (dequan.h:569-570)
                    if (d == d0) break;
                    --d;
                    D = saveD[d * 32 + lane];
                    if (HAS_F) F = saveF[d * 32 + lane];
                    c = cand[d];
                    continue;
                }
                if (!A.count_all && ((++poll & 63u) == 0) && *(volatile const unsigned long long*)A.best_key < idx) { outcome = 3; break; }
                if (A.node_budget && nodes > A.node_budget) { outcome = 2; break; }   // root probe: the host runs the split search instead
                const uint32_t Fx = HAS_F ? __shfl_sync(FULL, F, d) : 0u;
                if (d == nv - 1) {
                    // last variable: each remaining value is a node; valid ones are solutions, nothing left to filter
                    const uint32_t valid = c & ~Fx;
                    if (A.count_all) {
                        nodes += __popc(c);
                        if (valid && A.enum_out) enum_append(A, idx, nv, val, T.order, valid, sols, lane);
                        sols += __popc(valid);
                        if (valid && !have_first) {
                            have_first = true;
                            if (lane == 0) val[d] = (uint8_t)(__ffs(valid) - 1);
                            __syncwarp();
                            record_first();
                        }
                        c = 0;
                        continue;
                    }
                    if (valid) {
                        const int b = __ffs(valid) - 1;
                        nodes += __popc(c & ((2u << b) - 1u));
                        if (lane == 0) val[d] = (uint8_t)b;
                        __syncwarp();
                        sols = 1; outcome = 1; have_first = true;
                        break;
                    }
                    nodes += __popc(c);
                    c = 0;
                    continue;
                }
                const int b = __ffs(c) - 1;
                c &= c - 1;
                ++nodes;                                                               // AssignVar, dequan.h:416-423
                if (HAS_F && ((Fx >> b) & 1u)) continue;                               // ValidateVarConstraints fails, dequan.h:535-538
                uint32_t nd, nf;
                if (filter(d, b, D, F, nd, nf)) continue;                              // a wipe-out: nothing was written
                saveD[d * 32 + lane] = D;
                if (HAS_F) saveF[d * 32 + lane] = F;
                if (lane == 0) { cand[d] = c; val[d] = (uint8_t)b; }
                D = nd; F = nf;
                ++d;
                __syncwarp();
                c = __shfl_sync(FULL, D, d);
            }
            if (A.count_all && outcome != 2) outcome = sols ? 1 : 0;
        }
        if (outcome == 3) continue;                                                    // overtaken by an earlier prefix
        if (outcome == 2) { if (lane == 0) *A.gave_up = 1ull; break; }
        if (lane == 0) A.sub_nodes[idx] = nodes;
        if (A.count_all) { acc_nodes += nodes; acc_sols += sols; }
        else if (have_first) record_first();
        __syncwarp();
    }
    if (A.count_all && lane == 0) {
        atomicAdd(A.totals + 0, acc_sols);
        atomicAdd(A.totals + 1, acc_nodes);
    }
}

}  // namespace dq
