// dq_reg_graphs.cuh — register-resident warp engine for batches of k-colouring instances
// (BASELINE config C4: one graph per instance, variables AddIntVar(0,k), one
// OpConstraint(u, v, NotEqual, 0) per edge; k <= 4).
//
// One warp = one instance, and the whole search state is ONE 32-bit register per lane: lane j
// holds the domains of vertices j, j+32, j+64, ... as 4-bit fields (up to 8 x 32 = 256 vertices).
// The reference's forward check for "vertex x takes colour c"
// (OpConstraint::AplyArcConsistency -> Domain::Exclude on every unassigned neighbour, dequan.h:631-694, 985-1031)
// is, per lane,   D &= ~(later_neighbours_of_x_in_my_fields << c)   and the wipe-out test
// (dequan.h:663-668) is "does the result have a zero nibble" + one warp vote.  The static order is the
// vertex id (all domains have size k, Assignment::Reset ties by id, dequan.h:384-394), so "unassigned
// neighbour" = neighbour with a larger id, and the per-instance table peer[x][lane] (8 flag bits, one per
// field) holds only those: no assigned-or-not test at run time.
// A failing value is rejected BEFORE anything is written; a passing one saves the old register of exactly
// the lanes it changes (ballot + prefix popcount = the copy-on-first-write trail of
// Assignment::EnsureSavedDomain, dequan.h:442-452) and one packed word per level.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dq {

constexpr int kRegWarpsPerCta = 4;

struct RegGraphsArgs {
    int nv, k;
    const long long* edge_off;  // [n+1]
    const uint8_t* edges;       // [total][2]
    long long n;
    unsigned long long budget;
    unsigned long long* cursor;
    uint8_t* colours;           // [n][nv]
    unsigned long long* nodes;
    uint8_t* status;
    unsigned long long* totals; // [0]=sat [1]=unsat [2]=budget [3]=nodes
};

// per-warp shared memory: peer table (nv x 32 bytes), per-level words (nv x 8 bytes), trail (nv*k + 32 words)
__host__ __device__ inline size_t reg_graphs_warp_bytes(int nv, int k) {
    const size_t nvp = (size_t)((nv + 3) & ~3);
    return nvp * 32 + nvp * 8 + ((size_t)nv * k + 32) * 4;
}

// 8 flag bits -> one bit per 4-bit field
__device__ __forceinline__ uint32_t spread8(uint32_t x) {
    uint32_t t = (x | (x << 12)) & 0x000F000Fu;
    t = (t | (t << 6)) & 0x03030303u;
    return (t | (t << 3)) & 0x11111111u;
}

__global__ void __launch_bounds__(kRegWarpsPerCta * 32)
k_batch_graphs_reg(RegGraphsArgs A) {
    extern __shared__ __align__(16) unsigned char rg_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int nv = A.nv;
    const size_t nvp = (size_t)((nv + 3) & ~3);
    unsigned char* base = rg_raw + (size_t)wib * reg_graphs_warp_bytes(nv, A.k);
    uint32_t* peer_w = reinterpret_cast<uint32_t*>(base);                       // [nv][8] words = [nv][32] bytes
    const uint8_t* peer_b = base;
    unsigned long long* level = reinterpret_cast<unsigned long long*>(base + nvp * 32);   // [nv]
    uint32_t* trail = reinterpret_cast<uint32_t*>(base + nvp * 32 + nvp * 8);
    const uint32_t fullk = (1u << A.k) - 1u;
    unsigned long long t_sat = 0, t_unsat = 0, t_budget = 0, t_nodes = 0;

    for (;;) {
        long long i = 0;
        if (lane == 0) i = (long long)atomicAdd(A.cursor, 1ull);
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i >= A.n) break;

        // ---- peer table of this instance: for every edge, the later endpoint is a flag in the earlier one's row ----
        for (int w = lane; w < nv * 8; w += 32) peer_w[w] = 0;
        __syncwarp();
        const long long e0 = A.edge_off[i], e1 = A.edge_off[i + 1];
        for (long long e = e0 + lane; e < e1; e += 32) {
            const int u = A.edges[2 * e], v = A.edges[2 * e + 1];
            const int lo = min(u, v), hi = max(u, v);
            const int byte = lo * 32 + (hi & 31);
            atomicOr(&peer_w[byte >> 2], (1u << (hi >> 5)) << ((byte & 3) * 8));
        }
        __syncwarp();

        // ---- domains: full for real vertices, a non-empty dummy for the fields beyond nv ----
        uint32_t D = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) D |= ((lane + 32 * j < nv) ? fullk : 0xFu) << (4 * j);

        // ---- explicit-stack DFS (ForwardCheckingStep, dequan.h:494-571), level d = vertex d ----
        unsigned long long nodes = 0;
        int outcome = 0, d = 0, top = 0;
        uint32_t last_val = 0;
        uint32_t c = (__shfl_sync(0xFFFFFFFFu, D, 0)) & 0xFu;
        for (;;) {
            if (c == 0) {                                     // every value tried: return false (dequan.h:569-570)
                if (d == 0) break;
                --d;
                const unsigned long long L = level[d];
                const uint32_t mask = (uint32_t)L, mark = (uint32_t)(L >> 32) & 0xFFFF;
                if ((mask >> lane) & 1u) D = trail[mark + __popc(mask & lt)];     // RestoreSavedDomainStep, dequan.h:431-440
                top = (int)mark;
                c = (uint32_t)(L >> 48) & 0xFu;
                continue;
            }
            const uint32_t b = __ffs(c) - 1;
            c &= c - 1;
            ++nodes;                                          // AssignVar, dequan.h:416-423
            if (A.budget && nodes > A.budget) { outcome = 2; break; }
            if (d == nv - 1) { last_val = b; outcome = 1; break; }              // nothing left to filter: a solution
            const uint32_t pm = spread8(peer_b[d * 32 + lane]) << b;
            const uint32_t nd = D & ~pm;
            const bool wiped = ((nd - 0x11111111u) & ~nd & 0x88888888u) != 0;   // some later neighbour is left without a colour
            if (__any_sync(0xFFFFFFFFu, wiped)) continue;
            const bool ch = nd != D;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, ch);
            if (ch) trail[top + __popc(mask & lt)] = D;
            D = nd;
            if (lane == 0) level[d] = (unsigned long long)mask | ((unsigned long long)(uint32_t)top << 32) | ((unsigned long long)c << 48) | ((unsigned long long)b << 52);
            top += __popc(mask);
            ++d;
            __syncwarp();
            c = (__shfl_sync(0xFFFFFFFFu, D, d & 31) >> (4 * (d >> 5))) & 0xFu;
        }

        uint8_t* out = A.colours + (size_t)i * nv;
        for (int v = lane; v < nv; v += 32)
            out[v] = outcome == 1 ? (v == nv - 1 ? (uint8_t)last_val : (uint8_t)((level[v] >> 52) & 0xF)) : (uint8_t)0xFF;
        if (lane == 0) { A.nodes[i] = nodes; A.status[i] = (uint8_t)outcome; }
        t_nodes += nodes;
        t_sat += outcome == 1; t_unsat += outcome == 0; t_budget += outcome == 2;
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(A.totals + 0, t_sat); atomicAdd(A.totals + 1, t_unsat);
        atomicAdd(A.totals + 2, t_budget); atomicAdd(A.totals + 3, t_nodes);
    }
}

}  // namespace dq
