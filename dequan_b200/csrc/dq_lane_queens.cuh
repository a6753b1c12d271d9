// dq_lane_queens.cuh — lane-per-subtree forward-checking DFS for the N-Queens model class
// (CLASS_QUEENS: N variables on [0,N), NotEqual with offsets {0, +-(j-i)} for every pair —
// exactly /root/reference/test/main-test.cpp:36-49, recognised by the model compiler).
//
// Every lane owns one subtree and keeps its whole search state in registers:
//   A            values taken by assigned variables               (column mask)
//   L, R         the two diagonal masks as 64-bit words, shifted one bit per depth, so the current
//                domain of the variable at distance j+1 is   full & ~(A | L<<j | R>>j)
//                i.e. the closed form of what OpConstraint::AplyArcConsistency + Domain::Exclude
//                (dequan.h:631-694, 985-1031) leave in current_domains[d+1+j]
//   stk          chosen value per depth below the split, 5 bits each (the explicit DFS stack;
//                undo = shift the diagonals back and clear the popped bit, no trail memory)
// Forward checking of a candidate = test every later variable's domain for emptiness (the
// reference stops at the first wipe-out, dequan.h:514; so does the loop here).
// Node accounting is the reference's: one node per AssignVar call (dequan.h:416-423), i.e. per
// value of the current (filtered) domain of the next variable, whether or not its check succeeds.
//
// Work distribution, two kernels and no host round trip in between:
//   k_queens_items : item j is the base-N number whose k digits are the values of the first k
//                    variables.  One lane per item decodes it and replays the k assignments with
//                    forward checking; a surviving item is appended (warp-aggregated atomic) to a
//                    record list in HBM {key, A, L, R}.  A node above the split is counted by the
//                    single item that extends it with zeros, so the node total stays exact.
//   k_queens_lane  : persistent lanes pull records (two coalesced 16-byte loads per lane) and run
//                    the register-resident DFS below them.
// Items are dealt round-robin to partitions (multi-GPU): partition r owns items j = r (mod parts).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dq {

struct __align__(16) QueensRecord {      // one FC-surviving prefix = one subtree to search
    uint32_t key;                        // item index = DFS order of the prefix
    uint32_t a, llo, lhi, rlo, rhi;      // search state after the k prefix assignments
    uint32_t pad0, pad1;
};

struct QueensLaneArgs {
    int n, k;                         // board size; digits (split depth) per item
    unsigned long long n_items;       // n^k  (<= 2^27)
    unsigned int div_magic;           // ceil(2^32 / n): x / n == umulhi(x, magic) for x < 2^27, 2 <= n <= 32
    int part_rank, part_count;
    QueensRecord* records;            // [record_cap]
    unsigned long long record_cap;
    unsigned long long* n_records;    // valid items found (may exceed record_cap: then the host grows and reruns)
    unsigned long long* cursor;       // next record to search
    unsigned long long* totals;       // [0] solutions, [1] nodes
    unsigned long long* best_key;     // lowest item index that recorded a solution
    unsigned long long* sol_key;      // [n_threads]
    uint8_t* sol;                     // [n_threads][32] values by variable
};

constexpr int kQueensBlock = 256;

// Rare path: this lane just found the DFS-first solution of its item and the item may be the
// globally first one.  levels 0..k-1 come from the item digits, k..d-1 from the stack.
__device__ __noinline__ void queens_record_first(unsigned long long* best_key, unsigned long long* sol_key, uint8_t* sol,
                                                 int n, int k, unsigned int div_magic, unsigned long long key,
                                                 unsigned long long stk_lo, unsigned long long stk_hi, int d, int v_d,
                                                 int v_last) {
    const unsigned long long old = atomicMin(best_key, key);
    if (key >= old) return;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t* out = sol + tid * 32;
    uint32_t rem = (uint32_t)key;
    for (int t = k - 1; t >= 0; t--) {
        const uint32_t q = __umulhi(rem, div_magic);
        out[t] = (uint8_t)(rem - q * n);
        rem = q;
    }
    for (int lvl = d - 1; lvl >= k; lvl--) {        // top of stack = deepest level
        out[lvl] = (uint8_t)(stk_lo & 31);
        stk_lo = (stk_lo >> 5) | (stk_hi << 59);
        stk_hi >>= 5;
    }
    out[d] = (uint8_t)v_d;
    out[d + 1] = (uint8_t)v_last;
    __threadfence();
    sol_key[tid] = key;
}

// Phase A: validate items, count the nodes above the split, emit records.
__global__ void __launch_bounds__(kQueensBlock)
k_queens_items(QueensLaneArgs A) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int N = A.n, K = A.k;
    const uint32_t full = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
    const unsigned long long mine = (A.n_items + A.part_count - 1 - A.part_rank) / A.part_count;   // items of this partition
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long tot_nodes = 0;
    // uniform trip count per warp so the warp-aggregated append below stays convergent
    const unsigned long long first = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    for (unsigned long long base = first; base < mine; base += stride) {
        const unsigned long long ord = base + lane;
        bool valid = ord < mine;
        uint32_t a = 0, llo = 0, lhi = 0, rlo = 0, rhi = 0, nodes = 0;
        const unsigned long long key = ord * (unsigned long long)A.part_count + (unsigned long long)A.part_rank;
        if (valid) {
            // digits of the item, level 0 = most significant, packed 5 bits per level
            uint32_t rem = (uint32_t)key;
            unsigned long long dg = 0;
            for (int t = K - 1; t >= 0; t--) {
                const uint32_t q = __umulhi(rem, A.div_magic);
                dg |= (unsigned long long)(rem - q * N) << (5 * t);
                rem = q;
            }
            for (int i = 0; i < K; i++) {
                const uint32_t bit = 1u << ((dg >> (5 * i)) & 31);
                if (!(full & ~(a | llo | rhi) & bit)) { valid = false; break; }      // value not in the current domain: no node
                if ((dg >> (5 * (i + 1))) == 0) ++nodes;                             // this item is the node's representative
                const uint32_t lb = llo | bit, rb = rhi | bit;
                a |= bit;
                lhi = __funnelshift_l(lb, lhi, 1); llo = lb << 1;
                rlo = __funnelshift_r(rlo, rb, 1); rhi = rb >> 1;
                for (int j = 0; j <= N - 2 - i; j++)
                    if (((a | (llo << j) | (rhi >> j)) & full) == full) { valid = false; break; }   // wipe-out
                if (!valid) break;
            }
            tot_nodes += nodes;
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
        if (m) {
            unsigned long long slot = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) slot = atomicAdd(A.n_records, (unsigned long long)__popc(m));
            slot = __shfl_sync(0xFFFFFFFFu, slot, leader) + __popc(m & lt);
            if (valid && slot < A.record_cap) {
                uint4* dst = reinterpret_cast<uint4*>(A.records + slot);
                dst[0] = make_uint4((uint32_t)key, a, llo, lhi);
                dst[1] = make_uint4(rlo, rhi, 0u, 0u);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) tot_nodes += __shfl_down_sync(0xFFFFFFFFu, tot_nodes, o);
    if (lane == 0 && tot_nodes) atomicAdd(A.totals + 1, tot_nodes);
}

// Phase B: persistent lanes, register-resident DFS below each record.
template <bool STACK128>
__global__ void __launch_bounds__(kQueensBlock)
k_queens_lane(QueensLaneArgs A) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int N = A.n, K = A.k;
    const uint32_t full = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
    const unsigned long long n_found = *A.n_records;
    const unsigned long long n_rec = n_found < A.record_cap ? n_found : A.record_cap;   // overflow: the host grows the list and reruns
    unsigned long long tot_nodes = 0, tot_sols = 0;

    // per-lane search state
    uint32_t a = 0, llo = 0, lhi = 0, rlo = 0, rhi = 0, cand = 0, key = 0;
    unsigned long long stk = 0, stk_hi = 0;
    uint32_t nodes = 0, sols = 0;
    int d = 0;
    bool have = false, done = false, item_found = false;

    for (;;) {
        // ---- refill: lanes without a subtree take the next records ----
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !have && !done);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(A.cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!have && !done) {
                const unsigned long long r = base + __popc(need & lt);
                if (r >= n_rec) done = true;
                else {
                    const uint4* src = reinterpret_cast<const uint4*>(A.records + r);
                    const uint4 r0 = __ldg(src), r1 = __ldg(src + 1);
                    key = r0.x; a = r0.y; llo = r0.z; lhi = r0.w; rlo = r1.x; rhi = r1.y;
                    nodes = 0; sols = 0; item_found = false;
                    have = true;
                    d = K;
                    stk = 0; stk_hi = 0;
                    cand = full & ~(a | llo | rhi);
                }
            }
            if (__all_sync(0xFFFFFFFFu, done && !have)) break;
        }

        if (have) {
            if (cand == 0) {
                // ---- every value tried at this depth: return to the parent (dequan.h:569-570) ----
                if (d == K) {
                    have = false;
                    tot_nodes += nodes; tot_sols += sols;
                } else {
                    const uint32_t v = (uint32_t)stk & 31u;
                    if (STACK128) { stk = (stk >> 5) | (stk_hi << 59); stk_hi >>= 5; } else stk >>= 5;
                    const uint32_t bit = 1u << v;
                    --d;
                    a &= ~bit;
                    llo = __funnelshift_r(llo, lhi, 1) & ~bit; lhi >>= 1;
                    rhi = __funnelshift_l(rlo, rhi, 1) & ~bit; rlo <<= 1;
                    cand = full & ~(a | llo | rhi) & ~((bit << 1) - 1u);     // values after v, ascending order (dequan.h:554-562)
                }
            }
            if (have && cand) {
                // ---- AssignVar(next value) + forward check ----
                const uint32_t bit = cand & (0u - cand);
                cand ^= bit;
                ++nodes;
                const uint32_t lb = llo | bit, rb = rhi | bit;
                const uint32_t na = a | bit, nllo = lb << 1, nrhi = rb >> 1;
                bool ok = true;
                const int last = N - 2 - d;
                for (int j = 0; j <= last; j++)
                    if (((na | (nllo << j) | (nrhi >> j)) & full) == full) { ok = false; break; }
                if (ok) {
                    if (last == 0) {
                        // the child is the last variable: its whole domain is nodes, each one a solution
                        const uint32_t c = full & ~(na | nllo | nrhi);
                        const int pc = __popc(c);
                        nodes += pc;
                        if (!item_found) {
                            item_found = true;
                            if ((unsigned long long)key < *(volatile unsigned long long*)A.best_key)
                                queens_record_first(A.best_key, A.sol_key, A.sol, N, K, A.div_magic, key, stk, stk_hi, d, __ffs(bit) - 1, __ffs(c) - 1);
                        }
                        sols += pc;
                    } else {
                        const uint32_t v = __ffs(bit) - 1;
                        if (STACK128) { stk_hi = (stk_hi << 5) | (stk >> 59); }
                        stk = (stk << 5) | v;
                        a = na;
                        lhi = __funnelshift_l(lb, lhi, 1); llo = nllo;
                        rlo = __funnelshift_r(rlo, rb, 1); rhi = nrhi;
                        ++d;
                        cand = full & ~(a | llo | rhi);
                    }
                }
            }
        }
    }
    // warp-reduce the per-lane totals, one atomic pair per warp
    for (int o = 16; o > 0; o >>= 1) {
        tot_nodes += __shfl_down_sync(0xFFFFFFFFu, tot_nodes, o);
        tot_sols += __shfl_down_sync(0xFFFFFFFFu, tot_sols, o);
    }
    if (lane == 0) {
        atomicAdd(A.totals + 0, tot_sols);
        atomicAdd(A.totals + 1, tot_nodes);
    }
}

// Copies the solution recorded under `best_key` (if any) to out[0..n).
__global__ void k_queens_pick(const unsigned long long* __restrict__ best_key, const unsigned long long* __restrict__ sol_key,
                              const uint8_t* __restrict__ sol, size_t n_threads, int n, uint8_t* __restrict__ out) {
    const unsigned long long best = *best_key;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_threads; t += (size_t)gridDim.x * blockDim.x)
        if (sol_key[t] == best && best != 0xFFFFFFFFFFFFFFFFull)
            for (int i = 0; i < n; i++) out[i] = sol[t * 32 + i];
}

}  // namespace dq
