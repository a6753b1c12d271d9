// dq_group_graphs.cuh — lane-group engine for batches of k-colouring instances (BASELINE config C4: one graph per
// instance, variables AddIntVar(0,k), one OpConstraint(u, v, NotEqual, 0) per edge; k <= 4, <= 254 vertices).
//
// G lanes (1, 2, 4, 8, 16 or 32: the host picks by batch size) own ONE instance, so a warp searches 32/G
// instances at once and every instruction of the search loop advances all of them: the register-resident warp
// engine (dq_reg_graphs.cuh) spends ~100 warp instructions per node on a graph where a vertex has two or three
// later neighbours, i.e. 29 of its 32 lanes have nothing to filter.
//
// State of one instance, all in shared memory:
//   S[q]      one 32-bit word per vertex, byte c = "who took colour c away from q": 0 = still in q's domain, d+2 =
//             removed by the assignment of vertex d (OpConstraint::AplyArcConsistency -> Domain::Exclude,
//             /root/reference/dequan.h:631-694, 985-1031), 1 = q itself is assigned colour c, 0xFE = c >= k.
//             The current domain of q is the set of zero bytes; undoing the assignment of d
//             (Assignment::RestoreSavedDomainStep, dequan.h:431-440) clears exactly the bytes that hold d+2 — no
//             trail, no per-level record.
//   adj       deg[nvp] (later-neighbour count per vertex) followed by the later neighbours of vertex 0, 1, ...:
//             the static order is the vertex id (all domains have k values, Assignment::Reset ties by id,
//             dequan.h:384-394), so "unassigned neighbour" = neighbour with a larger id.  Built once per batch by
//             k_graphs_adjacency and brought in with ONE bulk copy (cp.async.bulk + mbarrier) per instance.
// One trip of the search loop handles a whole LEVEL: the forward check of every candidate colour of vertex d at
// once — colour c wipes out a later neighbour q exactly when q's domain is {c} (dequan.h:663-668), so one pass
// over the later neighbours yields the set F of failing colours; the first colour of the domain outside F is
// the one the reference descends with, the colours it tried before that one are nodes that failed their check
// (one AssignVar each, dequan.h:416-423), all counted.  A trip that returns to a level undoes the level's old
// assignment and retries with the colours above it in the same pass over the neighbours.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dq {

struct GroupGraphsArgs {
    int nv, k;
    int nvp;                    // nv rounded up to a multiple of 16
    int stride;                 // bytes of one adjacency record (multiple of 16): deg[nvp] + later neighbours
    const long long* edge_off;  // [n+1]                                  (k_graphs_adjacency)
    const uint8_t* edges;       // [total][2], 16-byte aligned, readable up to the next multiple of 16 bytes
    long long edge_bytes;       // 2 * total rounded up to 16
    uint8_t* adj;               // [n][stride]
    long long n;
    unsigned long long budget;
    unsigned long long* cursor;
    uint8_t* colours;           // [n][nv]
    unsigned long long* nodes;
    uint8_t* status;
    unsigned long long* totals; // [0]=sat [1]=unsat [2]=budget [3]=nodes [4]=instances the adjacency kernel refused
};

constexpr int kGroupWarpsPerCta = 4;
constexpr int kAdjWarpsPerCta = 4;

// ---- bulk-copy (TMA) + mbarrier primitives: SASS UBLKCP / SYNCS ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// Adjacency records.  One warp per instance, instances dealt round-robin; the raw edge list of the NEXT instance is
// in flight (bulk copy into the other staging buffer) while the current one is counted, scanned and scattered in
// shared memory; the finished record leaves with one bulk store.
__host__ __device__ inline size_t graphs_adj_warp_bytes(int nvp, int stride, int stage_cap) {
    return (size_t)2 * stage_cap + (size_t)stride + (size_t)nvp * 4 * 2 + 16;
}

__global__ void __launch_bounds__(kAdjWarpsPerCta * 32)
k_graphs_adjacency(GroupGraphsArgs A, int stage_cap) {
    extern __shared__ __align__(128) unsigned char ga_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nvp = A.nvp;
    unsigned char* base = ga_raw + (size_t)wib * ((graphs_adj_warp_bytes(nvp, A.stride, stage_cap) + 127) & ~(size_t)127);
    auto stage = [&](int b) { return base + (size_t)b * stage_cap; };
    unsigned char* rec = base + 2 * (size_t)stage_cap;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(rec + A.stride);
    uint32_t* pos = cnt + nvp;
    const uint32_t mbar0 = smem_u32(pos + nvp);
    if (lane == 0) { mbar_init(mbar0, 1); mbar_init(mbar0 + 8, 1); mbar_init_fence(); }
    __syncwarp();
    const long long total_warps = (long long)gridDim.x * kAdjWarpsPerCta;
    long long i = (long long)blockIdx.x * kAdjWarpsPerCta + wib;
    uint32_t parity = 0;                                 // bit b: phase of staging buffer b
    // the 16-byte aligned span of the edge array that holds instance i's pairs
    auto span = [&](long long inst, long long* a0, uint32_t* bytes, uint32_t* head) {
        const long long b0 = 2 * A.edge_off[inst], b1 = 2 * A.edge_off[inst + 1];
        *a0 = b0 & ~15ll;
        const long long a1 = min((b1 + 15) & ~15ll, A.edge_bytes);
        *bytes = b1 > b0 ? (uint32_t)(a1 - *a0) : 0u;
        *head = (uint32_t)(b0 - *a0);
    };
    auto issue = [&](long long inst, int b) {
        long long a0; uint32_t bytes, head;
        span(inst, &a0, &bytes, &head);
        if (bytes && bytes <= (uint32_t)stage_cap) {
            mbar_expect_tx(mbar0 + 8 * b, bytes);
            bulk_g2s(smem_u32(stage(b)), A.edges + a0, bytes, mbar0 + 8 * b);
        }
    };
    int b = 0;
    if (i < A.n && lane == 0) issue(i, 0);
    unsigned long long refused = 0;
    for (; i < A.n; i += total_warps, b ^= 1) {
        const long long nxt = i + total_warps;
        if (nxt < A.n && lane == 0) issue(nxt, b ^ 1);
        long long a0; uint32_t bytes, head;
        span(i, &a0, &bytes, &head);
        const uint32_t m = (uint32_t)(A.edge_off[i + 1] - A.edge_off[i]);
        bool bad = bytes > (uint32_t)stage_cap || (uint32_t)nvp + m > (uint32_t)A.stride;
        if (bytes && bytes <= (uint32_t)stage_cap) { mbar_wait(mbar0 + 8 * b, (parity >> b) & 1u); parity ^= 1u << b; }
        const unsigned char* e = stage(b) + head;
        for (int q = lane; q < nvp; q += 32) cnt[q] = 0;
        __syncwarp();
        if (!bad)
            for (uint32_t x = lane; x < m; x += 32) {
                const int u = e[2 * x], v = e[2 * x + 1];
                if (u >= A.nv || v >= A.nv || u == v) bad = true;      // (the host-buffer entry point has refused these already)
                else atomicAdd(&cnt[min(u, v)], 1u);
            }
        bad = __any_sync(0xFFFFFFFFu, bad);
        __syncwarp();
        // exclusive scan of the per-vertex counts: lane l owns vertices 8l .. 8l+7 (nvp <= 256)
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { const int q = lane * 8 + j; c[j] = q < nvp ? cnt[q] : 0u; sum += c[j]; }
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        uint32_t run = incl - sum;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int q = lane * 8 + j;
            if (q < nvp) { pos[q] = run; rec[q] = (unsigned char)c[j]; if (c[j] > 255u) bad = true; }
            run += c[j];
        }
        bad = __any_sync(0xFFFFFFFFu, bad);
        __syncwarp();
        if (!bad)
            for (uint32_t x = lane; x < m; x += 32) {
                const int u = e[2 * x], v = e[2 * x + 1];
                const uint32_t slot = atomicAdd(&pos[min(u, v)], 1u);
                rec[nvp + slot] = (unsigned char)max(u, v);
            }
        if (bad) {                                       // the search kernel sees an edgeless record and the status says why
            for (int q = lane; q < nvp; q += 32) rec[q] = 0;
            if (lane == 0) { A.status[i] = 0xFE; refused++; }
        }
        __syncwarp();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(A.adj + (size_t)i * A.stride, smem_u32(rec), (uint32_t)A.stride);
            bulk_commit();
            bulk_wait_read0();
        }
        __syncwarp();
    }
    if (lane == 0 && refused) atomicAdd(A.totals + 4, refused);
}

// ---------------------------------------------------------------------------------------------------------------
// bit 7 of every zero byte of w (exact, no borrow artefacts)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t w) { return ~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w | 0x7F7F7F7Fu); }
__device__ __forceinline__ uint32_t byte_of(uint32_t w, uint32_t c) { return __byte_perm(w, 0u, 0x4440u | c); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// per-warp shared memory of k_graphs_group: S words interleaved over the warp's instances, one adjacency record and
// one mbarrier per instance
__host__ __device__ inline size_t graphs_group_warp_bytes(int nvp, int stride, int g) {
    const int ipw = 32 / g;
    return (size_t)ipw * ((size_t)nvp * 4 + (size_t)stride + 16);
}

// OR of `fail` (bits 7, 15, 23, 31 only) over the G lanes of each group, every lane of the warp taking part: group g
// parks its four bits at 7-g, 15-g, 23-g, 31-g, so ONE warp-wide REDUX.OR serves all (at most 8) groups of the warp.
template <int G>
__device__ __forceinline__ uint32_t group_or_fail(uint32_t fail, int g) {
    if (G == 1) return fail;
    if (G == 2) return fail | __shfl_xor_sync(0xFFFFFFFFu, fail, 1);
    const uint32_t all = __reduce_or_sync(0xFFFFFFFFu, fail >> g);
    return (all << g) & 0x80808080u;
}

// The loop is warp-synchronous: every trip, each group that holds an instance handles one level of it; groups without
// one fetch the next instance of the batch; the warp leaves when the batch is empty and every group is done.
template <int G>
__global__ void __launch_bounds__(kGroupWarpsPerCta * 32)
k_graphs_group(GroupGraphsArgs A) {
    extern __shared__ __align__(128) unsigned char gg_raw[];
    constexpr int IPW = 32 / G;
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane / G, lg = lane % G;
    const int nv = A.nv, nvp = A.nvp;
    unsigned char* wbase = gg_raw + (size_t)wib * graphs_group_warp_bytes(nvp, A.stride, G);
    // S[q][instance of the warp]: lanes of different groups never share a bank on the same q
    const uint32_t s_base = smem_u32(wbase) + 4u * (uint32_t)g;
    const uint32_t adj_s = smem_u32(wbase + (size_t)IPW * nvp * 4 + (size_t)g * A.stride);
    const uint32_t mbar = smem_u32(wbase + (size_t)IPW * ((size_t)nvp * 4 + A.stride) + (size_t)g * 16);
    auto S = [&](uint32_t q) { return s_base + q * (4u * IPW); };
    if (lg == 0) { mbar_init(mbar, 1); mbar_init_fence(); }
    __syncwarp();
    uint32_t parity = 0;
    const uint32_t init_word = A.k >= 4 ? 0u : (0xFEFEFEFEu << (8 * A.k));
    const unsigned long long budget = A.budget ? A.budget : ~0ull;
    unsigned long long t_sat = 0, t_unsat = 0, t_budget = 0, t_nodes = 0;

    // state of the group's instance (identical in its G lanes)
    bool have = false, idle = false, ret = false;
    long long inst = 0;
    unsigned long long nodes = 0;
    int d = 0;
    uint32_t off = 0;

    for (;;) {
        const bool need = !have && !idle;
        if (__any_sync(FULL, need)) {
            long long i_new = 0;
            if (need && lg == 0) i_new = (long long)atomicAdd(A.cursor, 1ull);
            i_new = __shfl_sync(FULL, i_new, g * G);
            bool fresh = false;
            if (need) {
                if (i_new >= A.n) idle = true;
                else {
                    inst = i_new;
                    fresh = true;
                    if (lg == 0) {
                        mbar_expect_tx(mbar, (uint32_t)A.stride);
                        bulk_g2s(adj_s, A.adj + (size_t)inst * A.stride, (uint32_t)A.stride, mbar);
                    }
                    for (int q = lg; q < nvp; q += G) asm volatile("st.shared.u32 [%0], %1;" ::"r"(S(q)), "r"(init_word) : "memory");
                    nodes = 0; d = 0; off = 0; ret = false;
                }
            }
            if (fresh) {
                mbar_wait(mbar, parity);
                parity ^= 1;
                have = true;
                if (A.status[inst] == 0xFE) {             // k_graphs_adjacency could not build this instance's record
                    if (lg == 0) { A.nodes[inst] = 0; A.status[inst] = 3; }
                    for (int v = lg; v < nv; v += G) A.colours[(size_t)inst * nv + v] = 0xFF;
                    have = false;
                }
            }
            __syncwarp();
            if (__all_sync(FULL, idle)) break;
        }

        // ---- one trip = one level entered or re-entered (ForwardCheckingStep, dequan.h:494-571), level d = vertex d ----
        uint32_t fail = 0, cand = 0, c_old = 0, deg = 0, q0 = 0, w0 = FULL;
        const uint32_t undo_id = (uint32_t)d + 2u;
        const bool last = d == nv - 1;
        bool has0 = false;
        if (have) {
            uint32_t wd = lds32(S(d));
            if (ret) {
                // back at level d: its assignment (the byte that reads 1) is undone, the colours above it are left
                const uint32_t mk = zero_bytes(wd ^ 0x01010101u);
                c_old = (31u - (uint32_t)__clz((int)mk)) >> 3;
                wd &= ~(0xFFu << (8 * c_old));
                if (lg == 0) sts8(S(d) + c_old, 0u);
                cand = zero_bytes(wd) & ~(mk | (mk - 1u));
            } else cand = zero_bytes(wd);
            if (!last) {
                deg = lds8(adj_s + d);
                // first pass over the later neighbours: one per lane, kept in registers for the write-back
                has0 = (uint32_t)lg < deg;
                if (has0) {
                    q0 = lds8(adj_s + nvp + off + lg);
                    w0 = lds32(S(q0));
                    if (ret && byte_of(w0, c_old) == undo_id) { w0 &= ~(0xFFu << (8 * c_old)); sts8(S(q0) + c_old, 0u); }
                    const uint32_t t = zero_bytes(w0);
                    if (__popc(t) == 1) fail |= t;
                }
                for (uint32_t j = lg + G; j < deg; j += G) {
                    const uint32_t q = lds8(adj_s + nvp + off + j);
                    uint32_t w = lds32(S(q));
                    if (ret && byte_of(w, c_old) == undo_id) { w &= ~(0xFFu << (8 * c_old)); sts8(S(q) + c_old, 0u); }
                    const uint32_t t = zero_bytes(w);
                    if (__popc(t) == 1) fail |= t;
                }
            }
        }
        fail = group_or_fail<G>(fail, g);
        if (have) {
            // colour c wipes out a later neighbour exactly when that neighbour's domain is {c} (dequan.h:663-668): the
            // colours of `cand` up to the first one outside `fail` are the nodes the reference visits at this level
            const uint32_t pass = cand & ~fail;
            const uint32_t bit = pass & (0u - pass);
            const uint32_t tried = pass ? cand & (bit | (bit - 1u)) : cand;
            nodes += last ? 1u : (uint32_t)__popc(tried);
            int outcome = -1;
            uint32_t last_col = 0;
            if (nodes > budget) outcome = 2;
            else if (last) { outcome = 1; last_col = (31u - (uint32_t)__clz((int)(cand & (0u - cand)))) >> 3; }
            else if (pass == 0u) {
                if (d == 0) outcome = 0;
                else { --d; off -= lds8(adj_s + d); ret = true; }
            } else {
                const uint32_t c = (31u - (uint32_t)__clz((int)bit)) >> 3;
                if (has0 && (w0 & (0xFFu << (8 * c))) == 0u) sts8(S(q0) + c, undo_id);
                for (uint32_t j = lg + G; j < deg; j += G) {
                    const uint32_t q = lds8(adj_s + nvp + off + j);
                    if (lds8(S(q) + c) == 0u) sts8(S(q) + c, undo_id);
                }
                if (lg == 0) sts8(S(d) + c, 1u);
                off += deg;
                ++d;
                ret = false;
            }
            if (outcome >= 0) {
                if (outcome == 2) nodes = budget + 1;
                uint8_t* out = A.colours + (size_t)inst * nv;
                for (int v = lg; v < nv; v += G) {
                    uint32_t col = 0xFFu;
                    if (outcome == 1) {
                        const uint32_t mk = zero_bytes(lds32(S(v)) ^ 0x01010101u);
                        col = v == nv - 1 ? last_col : (31u - (uint32_t)__clz((int)mk)) >> 3;
                    }
                    out[v] = (uint8_t)col;
                }
                if (lg == 0) {
                    A.nodes[inst] = nodes;
                    A.status[inst] = (uint8_t)outcome;
                    t_nodes += nodes;
                    t_sat += outcome == 1; t_unsat += outcome == 0; t_budget += outcome == 2;
                }
                have = false;
            }
        }
        __syncwarp();
    }
    if (lg == 0) {
        if (t_sat) atomicAdd(A.totals + 0, t_sat);
        if (t_unsat) atomicAdd(A.totals + 1, t_unsat);
        if (t_budget) atomicAdd(A.totals + 2, t_budget);
        if (t_nodes) atomicAdd(A.totals + 3, t_nodes);
    }
}

}  // namespace dq
