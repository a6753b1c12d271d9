// dq_group_graphs.cuh — lane-per-instance engine for batches of k-colouring instances (BASELINE config C4: one graph
// per instance, variables AddIntVar(0,k), one OpConstraint(u, v, NotEqual, 0) per edge; k <= 4, <= 254 vertices).
//
// One LANE owns one instance, so a warp searches 32 instances at once and every instruction of the search loop
// advances all of them (the register-resident warp engine, dq_reg_graphs.cuh, spends ~100 warp instructions per
// node on a graph where a vertex has two or three later neighbours: 29 of its 32 lanes have nothing to filter).
// What keeps a lane fast is instruction-level parallelism: the eight neighbour slots of a level are eight
// INDEPENDENT load -> test chains, issued back to back.
//
// State of one instance, in shared memory, every array interleaved over the 32 lanes ([index][lane]: conflict-free):
//   S[q]      one 32-bit word per vertex, byte c = "who took colour c away from q": 0 = still in q's domain, 254-u =
//             removed by the assignment of vertex u (OpConstraint::AplyArcConsistency -> Domain::Exclude,
//             /root/reference/dequan.h:631-694, 985-1031), 255 = c >= k.  At level d a byte below 255-d is either 0 or the
//             mark of a vertex that is no longer assigned, so the current domain of q is the set of bytes < 255-d:
//             backtracking (Assignment::RestoreSavedDomainStep, dequan.h:431-440) stores NOTHING — stale marks are
//             ignored, and wiped when the vertex that could have left them is assigned again (its trip rewrites the
//             words of all its later neighbours).  S[nv] is a dummy (all 255) that the padding of the neighbour rows
//             points at: it never looks like a singleton and is never written.
//   rows[d]   8 bytes: the first eight later neighbours of vertex d, padded with nv; a vertex with more has 0xFF in the
//             last slot and the rest in a short (vertex, neighbour) list, walked by a cold path.  The static order
//             is the vertex id (all domains have k values, Assignment::Reset ties by id, dequan.h:384-394), so
//             "unassigned neighbour" = neighbour with a larger id.
//   co[i]     two bytes: col[i], the colour vertex i holds (written on the way down, read on the way back and for
//             the result), and open[i], the stack of the levels that still have untried colours: a failed level
//             returns straight to the top one — the levels in between have nothing left to try and, with no undo
//             to do, need no visit (they are four fifths of all returns on G(200, 4.2/199)).
// The per-instance records (rows + overflow list) are built once per batch by k_graphs_adjacency (bulk-copy staged
// edge lists in, bulk store out) and fetched by the search kernel with one bulk copy (cp.async.bulk + mbarrier) each.
// One trip of the search loop handles a whole LEVEL: the forward check of every candidate colour of vertex d at
// once — colour c wipes out a later neighbour q exactly when q's domain is {c} (dequan.h:663-668), so one pass
// over the later neighbours yields the set F of failing colours; the first colour of the domain outside F is
// the one the reference descends with, the colours it tried before that one are nodes that failed their check
// (one AssignVar each, dequan.h:416-423), all counted.  The trip is straight-line predicated code; everything rare —
// fetching the next instance, overflow neighbours, finishing an instance — sits behind one warp vote.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace dq {

struct GroupGraphsArgs {
    int nv, k;
    int nvp;                    // (nv + 16) & ~15: room for the dummy word S[nv]
    int rw;                     // neighbour slots per row (8)
    int over_cap;               // (vertex, neighbour) pairs the overflow list of one record can hold
    int over_smem;              // k_graphs_lane: how many of them a lane keeps in shared memory (the rest are read from the record)
    int stride;                 // bytes of one adjacency record (16 x an odd number)
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
    unsigned long long* totals; // [0]=sat [1]=unsat [2]=budget [3]=nodes [4]=instances the adjacency kernel refused [5]=longest overflow list
};

// record layout: rows[nvp][rw] | xflag[nvp] | u16 n_over, 14 bytes unused | over[over_cap][2]
__host__ __device__ inline int graphs_record_bytes(int nvp, int rw, int over_cap) {
    int b = nvp * rw + nvp + 16 + 2 * over_cap;
    b = (b + 15) & ~15;
    if (((b >> 4) & 1) == 0) b += 16;           // 16 x odd: the records of a warp's instances start in different banks
    return b;
}

constexpr int kGraphRow = 8;          // neighbour slots per vertex row
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
// in flight (bulk copy into the other staging buffer) while the current one is scattered into its rows in shared
// memory; the finished record leaves with one bulk store.
__host__ __device__ inline size_t graphs_adj_warp_bytes(int nvp, int stride, int stage_cap) {
    return (((size_t)2 * stage_cap + (size_t)stride + (size_t)nvp * 4 + 32) + 127) & ~(size_t)127;
}

// STATS_ONLY: the same pass over the edge lists, but all it produces is the largest overflow list any instance of
// the batch needs for rows of A.rw slots (totals[5], atomicMax) — the host sizes the records with it.
template <bool STATS_ONLY>
__global__ void __launch_bounds__(kAdjWarpsPerCta * 32)
k_graphs_adjacency(GroupGraphsArgs A, int stage_cap) {
    extern __shared__ __align__(128) unsigned char ga_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nv = A.nv, nvp = A.nvp, rw = A.rw;
    unsigned char* base = ga_raw + (size_t)wib * graphs_adj_warp_bytes(nvp, A.stride, stage_cap);
    auto stage = [&](int b) { return base + (size_t)b * stage_cap; };
    unsigned char* rec = base + 2 * (size_t)stage_cap;
    unsigned char* xflag = rec + (size_t)nvp * rw;
    unsigned char* hdr = xflag + nvp;
    unsigned char* over = hdr + 16;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(rec + A.stride);
    uint32_t* n_over = cnt + nvp;
    const uint32_t mbar0 = smem_u32(cnt + nvp + 2);
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
        bool bad = bytes > (uint32_t)stage_cap;
        if (bytes && bytes <= (uint32_t)stage_cap) { mbar_wait(mbar0 + 8 * b, (parity >> b) & 1u); parity ^= 1u << b; }
        const unsigned char* e = stage(b) + head;
        for (int q = lane; q < nvp; q += 32) { cnt[q] = 0; xflag[q] = 0; }
        if (!STATS_ONLY)
            for (int x = lane; x < nvp * rw; x += 32) rec[x] = (unsigned char)nv;      // padding: the dummy vertex
        if (lane == 0) *n_over = 0;
        __syncwarp();
        if (!bad)
            for (uint32_t x = lane; x < m; x += 32) {
                const int u = e[2 * x], v = e[2 * x + 1];
                if (u >= nv || v >= nv || u == v) { bad = true; continue; }     // (the host-buffer entry point has refused these already)
                const int lo = min(u, v), hi = max(u, v);
                const uint32_t slot = atomicAdd(&cnt[lo], 1u);
                if (slot < (uint32_t)rw) { if (!STATS_ONLY) rec[lo * rw + slot] = (unsigned char)hi; }
                else {
                    const uint32_t o = atomicAdd(n_over, 1u);
                    if (!STATS_ONLY && o < (uint32_t)A.over_cap) { over[2 * o] = (unsigned char)lo; over[2 * o + 1] = (unsigned char)hi; }
                }
            }
        __syncwarp();
        // a vertex with more neighbours than slots: its last slot becomes a marker (a byte above nv: the search kernel's
        // cue to walk the overflow list; which byte says where the vertex's pairs start in the list, sorted by vertex
        // below), the neighbour that sat there joins the list
        for (int q = lane; q < nvp; q += 32)
            if (cnt[q] > (uint32_t)rw) {
                const uint32_t o = atomicAdd(n_over, 1u);
                if (!STATS_ONLY) {
                    if (o < (uint32_t)A.over_cap) { over[2 * o] = (unsigned char)q; over[2 * o + 1] = rec[q * rw + rw - 1]; }
                    rec[q * rw + rw - 1] = 0xFF;
                    xflag[q] = 1;
                }
            }
        __syncwarp();
        if (!STATS_ONLY && !bad && *n_over && *n_over <= (uint32_t)A.over_cap) {
            const uint32_t no = *n_over;
            unsigned char* tmp = stage(b);                 // (the edge list has been consumed; 2 * no <= 2 * m bytes fit)
            for (int q = lane; q < nvp; q += 32)
                if (xflag[q]) {
                    uint32_t before = 0;
                    for (uint32_t x = 0; x < no; x++) before += over[2 * x] < (unsigned char)q;
                    cnt[q] = before;                       // cursor of the vertex's pairs in the sorted list
                    // marker 255 - start, kept above nv (a start that does not fit: the walk skips forward from the one that does)
                    rec[q * rw + rw - 1] = (unsigned char)(255u - min(before, (uint32_t)(254 - nv)));
                }
            __syncwarp();
            for (uint32_t x = lane; x < no; x += 32) {
                const uint32_t pos = atomicAdd(&cnt[over[2 * x]], 1u);
                tmp[2 * pos] = over[2 * x]; tmp[2 * pos + 1] = over[2 * x + 1];
            }
            __syncwarp();
            for (uint32_t x = lane; x < 2 * no; x += 32) over[x] = tmp[x];
            __syncwarp();
        }
        if (STATS_ONLY) {
            if (lane == 0 && *n_over) atomicMax(A.totals + 5, (unsigned long long)*n_over);
            __syncwarp();
            continue;
        }
        bad = __any_sync(0xFFFFFFFFu, bad) || *n_over > (uint32_t)A.over_cap;
        if (lane == 0) { hdr[0] = (unsigned char)(*n_over & 0xFF); hdr[1] = (unsigned char)(*n_over >> 8); }
        if (bad) {                                       // the search kernel skips the instance and the status says why
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
// bit 7 of every byte of x that is (unsigned) below the byte replicated in y; yl = y & 0x7F7F7F7F
__device__ __forceinline__ uint32_t bytes_below(uint32_t x, uint32_t y, uint32_t yl) {
    const uint32_t t = (x | 0x80808080u) - yl;                // per byte, no borrow across bytes: bit 7 = (low 7 bits of x >= those of y)
    return ((~x & y) | (~(x ^ y) & ~t)) & 0x80808080u;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t addr) { uint32_t v; asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& lo, uint32_t& hi) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t lo, uint32_t hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
    asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u16 [%0], t;\n\t}" ::"r"(addr), "r"(v) : "memory");
}

// shared memory of one warp (= one CTA) of k_graphs_lane.
// Per level (= vertex) and lane: the S word, and what the search keeps about the level — the colour it took and the
// stack of open levels (levels that still hold untried colours).  K4 (four colours): a byte each, next to S.
// Up to three colours (PACKED): byte 3 of the S words is free — the domain test never looks at it — and holds the open
// levels as a linked list (S[d].byte3 = the open level below d), and the colours go 16 to a word: 136 bytes per level and
// warp instead of 192, i.e. 8 warps per SM instead of 5 at 200 vertices.
// (At 200 vertices and three colours: 201 S words + 13 colour words + the overflow pairs = 27.4 KB per warp, the most
// that still lets 8 warps share an SM's 228 KB; a staging buffer for the record would cost the eighth.)
__host__ __device__ inline size_t graphs_lane_warp_bytes(int nv, int nvp, int over_smem, bool packed) {
    return (size_t)(nv + 1) * 128 + (packed ? (size_t)((nv + 15) / 16) * 128 : (size_t)nvp * 64) + (size_t)over_smem * 64 + 16;
}

template <bool PACKED>
__global__ void __launch_bounds__(32)
k_graphs_lane(GroupGraphsArgs A) {
    extern __shared__ __align__(128) unsigned char gl_raw[];
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    constexpr int kLeftCap = 0x7FFF0000;
    const uint32_t lane = threadIdx.x;
    const int nv = A.nv, nvp = A.nvp;
    const uint32_t base_s = smem_u32(gl_raw);
    const uint32_t s_s = base_s + 4u * lane;                                  // S[q]    at s_s + q * 128
    const uint32_t co_s = base_s + (uint32_t)(nv + 1) * 128u + (PACKED ? 4u : 2u) * lane;   // K4: co[i] at co_s + i * 64: byte 0 col[i], byte 1 open[i]; PACKED: colours of levels 16j .. 16j+15 at co_s + j * 128
    const uint32_t over_base = base_s + (uint32_t)(nv + 1) * 128u + (PACKED ? (uint32_t)((nv + 15) / 16) * 128u : (uint32_t)nvp * 64u);
    const uint32_t over_s = over_base + 2u * lane;                            // over[p] at over_s + p * 64: (vertex | neighbour << 8), sorted by vertex
    const uint16_t* over_g = reinterpret_cast<const uint16_t*>(A.adj + (size_t)nvp * 9u) + 8;   // ... and in the record, for p >= A.over_smem
    auto over_pair = [&](uint32_t p) -> uint32_t { return p < (uint32_t)A.over_smem ? lds16(over_s + p * 64u) : (uint32_t)__ldg(over_g + p); };
    // the neighbour rows stay in the record (HBM, L2- and L1-resident while the search is around that depth): 8 bytes per
    // level and lane, the next level's row fetched a trip ahead
    const uint2* rows_g = reinterpret_cast<const uint2*>(A.adj);
    uint2 row_cur = make_uint2(0u, 0u);                 // rows[d] when row_ok
    bool row_ok = false;
    // lanes without an instance run the trip on whatever their slots hold: start from zeros (vertex 0, no neighbours)
    for (uint32_t x = 4u * lane; x < (uint32_t)graphs_lane_warp_bytes(nv, nvp, A.over_smem, PACKED); x += 128u) sts32(base_s + x, 0u);
    __syncwarp();
    const uint32_t init_word = A.k >= 4 ? 0u : (FULL << (8 * A.k));
    const unsigned long long budget = A.budget ? A.budget : ~0ull;
    unsigned long long t_sat = 0, t_unsat = 0, t_budget = 0, t_nodes = 0;

    // state of the lane's instance
    bool have = false, idle = false, pend = false, ret = false;
    long long inst = 0;
    unsigned long long nodes_base = 0;                   // nodes counted before `left` was last set
    int left = 0, left0 = 0;                             // nodes the budget still allows (32-bit window of it) / its start value
    int code = 0;                                        // why the trip stopped the instance: bit 0 a solution, bit 1 tree exhausted (neither: `left` ran out)
    uint32_t d = 0, sp = 0, n_over = 0;
    uint32_t top = 0xFFu;                                // PACKED: the deepest open level (0xFF: none)
    // bytes of an S word that are colours (the domain test ignores the others)
    const uint32_t ymul = A.k >= 4 ? 0x01010101u : (0x01010101u & ((1u << (8 * A.k)) - 1u));

    for (;;) {
        // ---- one trip = one level entered or re-entered (ForwardCheckingStep, dequan.h:494-571), level d = vertex d ----
        const uint32_t wd = lds32(s_s + d * 128u);
        if (!row_ok) { row_cur = __ldg(rows_g + d); row_ok = true; }
        const uint2 row_next = __ldg(rows_g + d + 1u);
        uint32_t r_lo = row_cur.x, r_hi = row_cur.y;
        uint32_t c_old, d_pop, cw = 0;                               // colour taken here before (meaningful when ret); level to go back to
        if constexpr (PACKED) {
            cw = lds32(co_s + (d >> 4) * 128u);
            c_old = (cw >> (2u * (d & 15u))) & 3u;
            if (ret) top = wd >> 24;                                 // back at an open level: it leaves the list (and rejoins below if colours are left)
            d_pop = top;
        } else {
            c_old = lds8(co_s + d * 64u);
            d_pop = lds8(co_s + sp * 64u - 63u);                     // open[sp - 1] (meaningful when sp > 0)
        }
        // the row of the level a dead end here goes back to, a trip ahead like the next level's (half of all trips end
        // that way, and a row loaded on arrival leaves its L2 latency exposed)
        const uint2 row_back = __ldg(rows_g + min(d_pop, (uint32_t)nv));
        const bool more = (r_hi >> 24) > (uint32_t)nv;               // a marker: the row goes on in the overflow list ...
        const uint32_t over_start = 255u - (r_hi >> 24);             // ... from this position on (the list is sorted by vertex)
        if (more) r_hi = (r_hi & 0x00FFFFFFu) | ((uint32_t)nv << 24);   // (the marker is not a vertex: look at the dummy instead)
        const uint32_t y = (255u - d) * ymul, yl = y & 0x7F7F7F7Fu;           // bytes below 255-d: in the domain
        const uint32_t mark = 254u - d;
        uint32_t a[8], w[8], f[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t q = ((j < 4 ? r_lo : r_hi) >> (8 * (j & 3))) & 0xFFu;
            a[j] = s_s + q * 128u;
            w[j] = lds32(a[j]);
        }
        uint32_t fail = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            f[j] = bytes_below(w[j], y, yl);
            fail |= (f[j] & (f[j] - 1u)) ? 0u : f[j];                // a neighbour down to one colour forbids it
        }
        // cold work pending in this warp?  (a lane without an instance, an instance that just ended, overflow neighbours)
        if (__any_sync(FULL, pend || (!have && !idle) || (have && more))) {
            if (__any_sync(FULL, pend || (!have && !idle))) {
                // ---- results out ----
                if (pend) {
                    unsigned long long nodes = nodes_base + (unsigned long long)(left0 - left);
                    if (code == 0 && nodes <= budget) {            // only the 32-bit window ran out: open the next one
                        nodes_base = nodes;
                        left0 = left = (int)min(budget - nodes, (unsigned long long)kLeftCap);
                        have = true;
                    } else {
                        const int outcome = nodes > budget ? 2 : ((code & 1) ? 1 : 0);
                        if (outcome == 2) nodes = budget + 1;
                        uint8_t* out = A.colours + (size_t)inst * nv;
                        for (int v = 0; v < nv; v++) {
                            uint32_t cv;
                            if constexpr (PACKED) cv = (lds32(co_s + ((uint32_t)v >> 4) * 128u) >> (2u * ((uint32_t)v & 15u))) & 3u;
                            else cv = lds8(co_s + (uint32_t)v * 64u);
                            out[v] = outcome == 1 ? (uint8_t)cv : (uint8_t)0xFF;
                        }
                        A.nodes[inst] = nodes;
                        A.status[inst] = (uint8_t)outcome;
                        t_nodes += nodes;
                        t_sat += outcome == 1; t_unsat += outcome == 0; t_budget += outcome == 2;
                        d = 0; sp = 0; top = 0xFFu; ret = false; row_ok = false;
                    }
                    pend = false;
                }
                // ---- next instances in ----
                uint32_t needy = __ballot_sync(FULL, !have && !idle);
                while (needy) {
                    const int t = __ffs((int)needy) - 1;
                    needy &= needy - 1u;
                    long long i_new = 0;
                    if (lane == 0) i_new = (long long)atomicAdd(A.cursor, 1ull);
                    i_new = __shfl_sync(FULL, i_new, 0);
                    if (i_new >= A.n) {                            // the batch is empty: this lane and the remaining ones are done
                        if ((int)lane == t || ((needy >> lane) & 1u)) idle = true;
                        break;
                    }
                    const bool refused = A.status[i_new] == 0xFE;  // k_graphs_adjacency could not build this instance's record
                    if (refused) {
                        if (lane == 0) { A.nodes[i_new] = 0; A.status[i_new] = 3; }
                        for (int v = lane; v < nv; v += 32) A.colours[(size_t)i_new * nv + v] = 0xFF;
                        needy |= 1u << t;                          // the lane still needs an instance
                    } else {
                        // the lane's columns: the state words, and the few neighbour pairs its rows had no room for
                        // (straight from the record: its rows are read from there level by level anyway)
                        for (uint32_t q = lane; q <= (uint32_t)nv; q += 32)
                            sts32(base_s + q * 128u + 4u * t, q == (uint32_t)nv ? FULL : init_word);
                        const uint16_t* tail = reinterpret_cast<const uint16_t*>(A.adj + (size_t)i_new * A.stride + (size_t)nvp * 9u);
                        const uint32_t n_ov = __ldg(tail);
                        for (uint32_t p = lane; p < (uint32_t)A.over_smem; p += 32)
                            sts16(over_base + p * 64u + 2u * t, p < n_ov ? (uint32_t)__ldg(tail + 8 + p) : 0xFFFFu);
                        if ((int)lane == t) {
                            inst = i_new; have = true; n_over = n_ov;
                            rows_g = reinterpret_cast<const uint2*>(A.adj + (size_t)i_new * A.stride);
                            over_g = tail + 8;
                            row_ok = false;
                            nodes_base = 0; d = 0; sp = 0; top = 0xFFu; ret = false;
                            left0 = left = (int)min(budget, (unsigned long long)kLeftCap);
                        }
                    }
                    __syncwarp();
                }
                if (__all_sync(FULL, idle)) break;
                continue;                                          // (the loads above are stale: start the trip over)
            }
            // ---- the neighbours of d beyond its row ----
            if (have && more) {
                uint32_t p = over_start;
                while (p < n_over && (over_pair(p) & 0xFFu) < d) ++p;
                for (; p < n_over; p++) {
                    const uint32_t pr = over_pair(p);
                    if ((pr & 0xFFu) != d) break;
                    const uint32_t fo = bytes_below(lds32(s_s + (pr >> 8) * 128u), y, yl);
                    if ((fo & (fo - 1u)) == 0u) fail |= fo;
                }
            }
        }
        // colour c wipes out a later neighbour exactly when that neighbour's domain is {c} (dequan.h:663-668): the
        // colours of `cand` up to the first one outside `fail` are the nodes the reference visits at this level
        uint32_t cand = bytes_below(wd, y, yl);
        if (ret) cand &= ~((0x100u << (8u * c_old)) - 1u);       // back at this level: the colours above the old one are left
        const uint32_t pass = cand & ~fail;
        const uint32_t bit = pass & (0u - pass);
        const bool descend = pass != 0u;
        const bool last = d == (uint32_t)(nv - 1);
        const uint32_t upto = bit | (bit - 1u);
        const uint32_t tried = descend ? cand & upto : cand;
        const uint32_t n_tried = last ? 1u : (((tried >> 7) * 0x01010101u) >> 24);   // the last variable's first value completes a solution
        const uint32_t unit = bit >> 7;                          // 1 << 8c
        const uint32_t put = unit * mark;
        const bool act = have && descend;
        // the neighbours' words: stale marks out, this vertex's mark in (where the colour was still there)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t m = (f[j] >> 7) * 255u;
            if (act && f[j] != 0u) sts32(a[j], (w[j] & ~m) | (put & m));
        }
        const bool rest = (cand & ~upto) != 0u;
        if (act) {
            if constexpr (PACKED) {
                const uint32_t sh = 2u * (d & 15u);
                sts32(co_s + (d >> 4) * 128u, (cw & ~(3u << sh)) | (((unit * 0x00010203u) >> 24) << sh));
                if (rest) sts8(s_s + d * 128u + 3u, top);            // joins the open levels
            } else {
                sts8(co_s + d * 64u, (unit * 0x00010203u) >> 24);
                sts8(co_s + sp * 64u + 1u, d);
            }
            if (more) {
                uint32_t p = over_start;
                while (p < n_over && (over_pair(p) & 0xFFu) < d) ++p;
                for (; p < n_over; p++) {
                    const uint32_t pr = over_pair(p);
                    if ((pr & 0xFFu) != d) break;
                    const uint32_t ao = s_s + (pr >> 8) * 128u;
                    const uint32_t wo = lds32(ao);
                    const uint32_t m = (bytes_below(wo, y, yl) >> 7) * 255u;
                    sts32(ao, (wo & ~m) | (put & m));
                }
            }
        }
        // down, back to the deepest level with colours left (return false, dequan.h:569-570), or done
        if (have) {
            left -= (int)n_tried;
            const bool sat = descend && last, unsat = !descend && (PACKED ? top == 0xFFu : sp == 0u);
            code = (sat ? 1 : 0) | (unsat ? 2 : 0);
            if constexpr (PACKED) { if (descend && rest) top = d; }
            else sp = descend ? sp + (rest ? 1u : 0u) : sp - 1u;
            d = descend ? d + 1u : d_pop;
            ret = !descend;
            row_ok = true;
            row_cur = descend ? row_next : row_back;
            if (sat || unsat || left < 0) { have = false; pend = true; if (unsat) { d = 0; sp = 0; top = 0xFFu; row_ok = false; } }
        }
    }
    if (t_sat) atomicAdd(A.totals + 0, t_sat);
    if (t_unsat) atomicAdd(A.totals + 1, t_unsat);
    if (t_budget) atomicAdd(A.totals + 2, t_budget);
    if (t_nodes) atomicAdd(A.totals + 3, t_nodes);
}

}  // namespace dq
