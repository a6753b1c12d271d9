// dq_lane_sudoku.cuh — lane-per-instance forward-checking DFS for batches of 9x9 Sudoku
// (CLASS_SUDOKU9: 81 variables on 1..9, NotEqual/AllDifferent over the 27 rows, columns and boxes
// — BASELINE config C3, recognised by the model compiler whatever mix of binary != and
// AllDifferent rows produced it).
//
// What the reference does per node (CSP::ForwardCheckingStep dequan.h:494-571 with
// OpConstraint/AllDifferent::AplyArcConsistency 631-694, 915-939 and Domain::Exclude 985-1031)
// collapses, for this model class, to a closed form: the current domain of an unassigned cell is
//     {1..9} minus the values assigned in its row, its column and its box
// so the whole search state of one instance is 27 nine-bit "used" masks.  Undo on backtrack
// (RestoreSavedDomainStep, dequan.h:431-440) is clearing one bit in three masks; there is no trail.
//
// Static order (Assignment::Reset, dequan.h:376-394): givens first (domain size 1) by id, then the
// blanks by id.  A node is an AssignVar call (dequan.h:416-423): one per given, then one per value
// tried at a blank.  The forward check after assigning value v at cell p asks whether any LATER
// blank peer q is left with no value: used(q) | v == all nine.  Cells are packed three to a word
// (10-bit fields, one word per (row, stack) = one box-row), so a check is one OR/AND/ADD per word:
// adding 1 to every field carries into the field's spare bit exactly when the field is all ones.
//
// Layout: every lane owns one instance; its masks, blank-field masks and DFS stack live in shared
// memory as [word][thread], which is bank-conflict-free whatever each lane indexes.
//
// Work distribution.  Node counts per puzzle are heavy-tailed (30 givens: median 363, max > 1e6), and a
// lane is slow, so long searches are cut into TASKS: a lane that has spent its node budget on a
// task stops, and hands every untried candidate set left on its stack to the task pool as an
// independent subtree ("piece"), each with a sub-interval of the parent's 64-bit DFS-order key
// range.  Pieces are searched by any lane in a later round.  Exactness of the per-puzzle node count
// (the reference's sequential count up to its first solution) comes from the keys: the first
// solution in DFS order is the found piece with the smallest key, and the puzzle's node count is
// the sum over its tasks with key <= that key.  Tasks with larger keys are speculative; they are
// skipped or abandoned as soon as a smaller-keyed solution is known.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dq {

constexpr int kSudokuBlock = 128;               // threads per CTA
constexpr uint32_t SK_ONES = 0x00100401u;       // 1 in each of the three 10-bit fields
constexpr uint32_t SK_SPARE = 0x20080200u;      // bit 9 of each field
constexpr uint32_t SK_FULL3 = 0x1FF7FDFFu;      // 0x1FF in each field
constexpr unsigned long long SK_KEY_END = 0x7FFFFFFFFFFFFFFFull;
constexpr uint8_t SK_STATUS_SPLIT = 0xFD;       // interim: the puzzle's tasks are still being accounted
constexpr uint8_t SK_STATUS_DEFER = 0xFE;       // interim: handed to the generic warp engine

// 48-byte digest of one instance, written by k_sudoku_digest and read by the lanes.
//   w[0..2]  rows' used masks, three rows per word (10-bit fields)
//   w[3..5]  columns' used masks, three columns per word
//   w[6..8]  boxes' used masks, three boxes (one band) per word
//   w[9..11] blank bitmap by cell id (81 bits); w[11] bit 31 = deferred to the generic engine
struct SudokuDigest { uint32_t w[12]; };

struct SudokuPiece {           // 32 bytes: one part of a split task's remaining search
    uint32_t puzzle;
    uint32_t snap_id;          // which stack snapshot it resumes from
    uint32_t levels;           // lo | hi << 8: the snapshot's stack levels [lo, hi] whose untried values it owns;
                               // 0xFFFFFFFF = null piece (pool overflow filler)
    uint32_t pad;
    unsigned long long key_lo, key_hi;
};
constexpr int kSnapWords = 12;                  // uint4 per stack snapshot: 96 u16 entries (untried | value << 9) per level

struct SudokuArgs {
    const SudokuDigest* digest;
    long long n;                           // instances
    int stride;
    const uint8_t* cells;                  // [n][stride] givens (for writing solutions)
    uint8_t* solution;                     // [n][stride]
    unsigned long long* nodes;             // [n]
    uint8_t* status;                       // [n]
    unsigned long long* best_key;          // [n] smallest key of a task that found a solution
    // task pool
    SudokuPiece* pieces;  unsigned long long piece_cap;
    uint4* snaps;         unsigned long long snap_cap;       // kSnapWords x uint4 per snapshot
    unsigned long long* piece_nodes;       // [piece_cap]
    uint8_t* piece_sol;                    // [piece_cap][81], valid where piece_found
    uint8_t* piece_found;                  // [piece_cap]
    // control block: [0] cursor  [1] piece tail  [2] snapshot tail  [3] round begin  [4] round end
    unsigned long long* ctrl;
    unsigned budget;                       // nodes per task before it is split
    unsigned long long user_budget;        // per-instance node budget of the API (0 = none)
    int round;                             // 0: fresh instances, >0: pool pieces [ctrl[3], ctrl[4])
};

// ---------------------------------------------------------------------------------------------
// Pass 0: one thread per instance, inputs staged through shared memory so that HBM reads coalesce.
__global__ void __launch_bounds__(128)
k_sudoku_digest(const uint8_t* __restrict__ cells, long long n, int stride, SudokuDigest* __restrict__ out,
                unsigned long long* __restrict__ best_key) {
    extern __shared__ uint8_t stage[];                       // [128][81]
    const long long first = (long long)blockIdx.x * 128;
    const int here = (int)min((long long)128, n - first);
    if (stride == 81) {
        const uint8_t* src = cells + first * 81;
        for (int i = threadIdx.x; i < here * 81; i += 128) stage[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < here * 81; i += 128) stage[i] = cells[(first + i / 81) * stride + i % 81];
    }
    __syncthreads();
    if ((int)threadIdx.x >= here) return;
    const uint8_t* g = stage + threadIdx.x * 81;
    uint32_t row[9], col[9], box[9], blank[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; i++) { row[i] = 0; col[i] = 0; box[i] = 0; }
    bool defer = false;
#pragma unroll
    for (int r = 0; r < 9; r++)
#pragma unroll
        for (int c = 0; c < 9; c++) {
            const int p = r * 9 + c;
            const uint32_t v = g[p];
            if (v == 0) { blank[p >> 5] |= 1u << (p & 31); continue; }
            if (v > 9) { defer = true; continue; }           // not a value of the template domain
            const uint32_t bit = 1u << (v - 1);
            const int b = (r / 3) * 3 + c / 3;
            if ((row[r] | col[c] | box[b]) & bit) defer = true;   // two givens clash: the exact node count is the generic engine's job
            row[r] |= bit; col[c] |= bit; box[b] |= bit;
        }
#pragma unroll
    for (int r = 0; r < 9; r++)
#pragma unroll
        for (int c = 0; c < 9; c++) {
            const int p = r * 9 + c;
            if (((blank[p >> 5] >> (p & 31)) & 1u) && (row[r] | col[c] | box[(r / 3) * 3 + c / 3]) == 0x1FFu) defer = true;  // a blank wiped out by the givens
        }
    SudokuDigest d;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        d.w[t] = row[3 * t] | (row[3 * t + 1] << 10) | (row[3 * t + 2] << 20);
        d.w[3 + t] = col[3 * t] | (col[3 * t + 1] << 10) | (col[3 * t + 2] << 20);
        d.w[6 + t] = box[3 * t] | (box[3 * t + 1] << 10) | (box[3 * t + 2] << 20);
        d.w[9 + t] = blank[t];
    }
    if (defer) d.w[11] |= 0x80000000u;
    uint4* o = reinterpret_cast<uint4*>(out + first + threadIdx.x);
    o[0] = make_uint4(d.w[0], d.w[1], d.w[2], d.w[3]);
    o[1] = make_uint4(d.w[4], d.w[5], d.w[6], d.w[7]);
    o[2] = make_uint4(d.w[8], d.w[9], d.w[10], d.w[11]);
    best_key[first + threadIdx.x] = 0xFFFFFFFFFFFFFFFFull;
}

// ---------------------------------------------------------------------------------------------
// Per-lane shared-memory state, all arrays [word][thread].
struct SudokuSmem {
    uint32_t rowr[9][kSudokuBlock];     // row used mask, replicated in the three fields
    uint32_t colp[3][kSudokuBlock];     // column used masks of one stack, one per field
    uint32_t boxr[9][kSudokuBlock];     // box used mask, replicated
    uint32_t blk[27][kSudokuBlock];     // (row, stack) -> 0x1FF in the fields of blank cells
    uint16_t stk[81][kSudokuBlock];     // per level: untried candidates | value index << 9
};

__device__ __forceinline__ uint32_t sk_rep(uint32_t x9) { return x9 * SK_ONES; }

struct SkCell { int r, s, f, band, box; };
__device__ __forceinline__ SkCell sk_decode(int p) {
    SkCell c;
    c.r = (p * 57) >> 9;                 // p / 9 for p < 81
    const int col = p - 9 * c.r;
    c.s = (col * 11) >> 5;               // col / 3 for col < 9
    c.f = col - 3 * c.s;
    c.band = (c.r * 11) >> 5;            // r / 3
    c.box = c.band * 3 + c.s;
    return c;
}

// next / previous blank cell in id order from the 81-bit bitmap (b0: cells 0-31, b1: 32-63, b2: 64-80)
__device__ __forceinline__ int sk_next_blank(uint32_t b0, uint32_t b1, uint32_t b2, int p) {
    // p in [-1, 80]; returns 81 if none
    const int q = p + 1;
    const uint32_t m0 = q < 32 ? (b0 >> q) << q : 0u;
    const uint32_t m1 = q < 32 ? b1 : (q < 64 ? (b1 >> (q - 32)) << (q - 32) : 0u);
    const uint32_t m2 = q < 64 ? b2 : (b2 >> (q - 64)) << (q - 64);
    if (m0) return __ffs(m0) - 1;
    if (m1) return 32 + __ffs(m1) - 1;
    if (m2) return 64 + __ffs(m2) - 1;
    return 81;
}
__device__ __forceinline__ int sk_prev_blank(uint32_t b0, uint32_t b1, uint32_t b2, int p) {
    // largest blank id < p (p in [0,81]); -1 if none
    const uint32_t m2 = p > 64 ? b2 & ((1u << (p - 64)) - 1u) : 0u;
    const uint32_t m1 = p >= 64 ? b1 : (p > 32 ? b1 & ((1u << (p - 32)) - 1u) : 0u);
    const uint32_t m0 = p >= 32 ? b0 : (p > 0 ? b0 & ((1u << p) - 1u) : 0u);
    if (m2) return 64 + 31 - __clz(m2);
    if (m1) return 32 + 31 - __clz(m1);
    if (m0) return 31 - __clz(m0);
    return -1;
}

// The search kernel: one launch per round.  Round 0 walks the fresh instances; later rounds walk
// the pieces the previous round produced.
__global__ void __launch_bounds__(kSudokuBlock)
k_sudoku_lane(SudokuArgs A) {
    extern __shared__ __align__(16) unsigned char sk_raw[];
    SudokuSmem& S = *reinterpret_cast<SudokuSmem*>(sk_raw);
    const int t = threadIdx.x;
    const int lane = t & 31;
    const uint32_t lt = (1u << lane) - 1u;

    const unsigned long long q_begin = A.round == 0 ? 0ull : A.ctrl[3];
    const unsigned long long q_end = A.round == 0 ? (unsigned long long)A.n : A.ctrl[4];
    const unsigned long long total_warps = (unsigned long long)gridDim.x * (kSudokuBlock / 32);

    // per-lane task state
    bool have = false, done = false;
    unsigned long long task = 0;            // round 0: instance id; else piece id
    uint32_t puzzle = 0;
    uint32_t b0 = 0, b1 = 0, b2 = 0;        // blank bitmap
    int p = 0;                              // current cell
    int sp = 0, base_sp = 0, nblank = 0;    // stack level of the current cell, the task's root level, blanks in the puzzle
    uint32_t cand = 0;                      // untried values at the current cell
    uint32_t nodes = 0, limit = 0;          // nodes tried by this task; split / stop threshold
    unsigned long long nodes_hi = 0;        // overflow of `nodes` for unlimited tasks
    bool stop_is_budget = false;            // reaching `limit` means the API node budget ran out (no split)
    unsigned long long key_lo = 0, key_hi = 0;
    unsigned poll = 0;
    // warp-uniform queue state
    unsigned long long chunk_pos = 0, chunk_end = 0;
    bool exhausted = false;

    for (;;) {
        // ---------------- refill ----------------
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !have && !done);
        if (need) {
            const uint32_t n_need = __popc(need);
            if (chunk_pos >= chunk_end && !exhausted) {
                unsigned long long base = 0;
                uint32_t size = 0;
                if (lane == 0) {
                    const unsigned long long cur = *(volatile unsigned long long*)(A.ctrl + 0);
                    const unsigned long long at = q_begin + cur;
                    const unsigned long long remaining = at < q_end ? q_end - at : 0;
                    size = (uint32_t)min(max(remaining / (8ull * total_warps), (unsigned long long)n_need), 64ull);
                    base = q_begin + atomicAdd(A.ctrl + 0, (unsigned long long)size);
                }
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                size = __shfl_sync(0xFFFFFFFFu, size, 0);
                chunk_pos = base;
                chunk_end = min(base + size, q_end);
                if (base >= q_end) { exhausted = true; chunk_end = chunk_pos; }
            }
            const unsigned long long avail = chunk_end - chunk_pos;
            if (!have && !done) {
                const uint32_t rank = __popc(need & lt);
                if (rank < avail) {
                    task = chunk_pos + rank;
                    // ---- load the task ----
                    uint32_t levels = 0, snap_id = 0;
                    if (A.round == 0) {
                        puzzle = (uint32_t)task;
                        key_lo = 0; key_hi = SK_KEY_END;
                    } else {
                        const uint4* pr = reinterpret_cast<const uint4*>(A.pieces + task);
                        const uint4 x = pr[0], y = pr[1];
                        puzzle = x.x; snap_id = x.y; levels = x.z;
                        key_lo = (unsigned long long)y.x | ((unsigned long long)y.y << 32);
                        key_hi = (unsigned long long)y.z | ((unsigned long long)y.w << 32);
                    }
                    const uint4* dg = reinterpret_cast<const uint4*>(A.digest + puzzle);
                    const uint4 d0 = __ldg(dg), d1 = __ldg(dg + 1), d2 = __ldg(dg + 2);
                    const uint32_t w[12] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
                    b0 = w[9]; b1 = w[10]; b2 = w[11] & 0x0001FFFFu;
                    nblank = __popc(b0) + __popc(b1) + __popc(b2);
                    have = true;
                    nodes = 0; nodes_hi = 0; poll = 0;
                    stop_is_budget = false;
                    limit = A.budget;
                    bool skip = false;
                    if (A.round == 0) {
                        if (w[11] & 0x80000000u) { A.status[puzzle] = SK_STATUS_DEFER; skip = true; }
                        else if (A.user_budget) {
                            const unsigned long long givens = 81 - nblank;
                            if (A.user_budget < givens) {        // the budget runs out among the givens
                                A.nodes[puzzle] = A.user_budget + 1; A.status[puzzle] = 2; skip = true;
                                uint8_t* out = A.solution + (size_t)puzzle * A.stride;
                                for (int i = 0; i < 81; i++) out[i] = 0;
                            } else if (A.user_budget - givens + 1 <= (unsigned long long)limit) {
                                limit = (uint32_t)(A.user_budget - givens + 1);
                                stop_is_budget = true;
                            }
                        }
                    } else if (levels == 0xFFFFFFFFu || *(volatile unsigned long long*)(A.best_key + puzzle) < key_lo) {
                        A.piece_nodes[task] = 0; A.piece_found[task] = 0; skip = true;   // null, or an earlier subtree already holds a solution
                    }
                    if (skip) have = false;
                    else {
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            S.rowr[3 * i][t] = sk_rep(w[i] & 0x1FF);
                            S.rowr[3 * i + 1][t] = sk_rep((w[i] >> 10) & 0x1FF);
                            S.rowr[3 * i + 2][t] = sk_rep((w[i] >> 20) & 0x1FF);
                            S.colp[i][t] = w[3 + i];
                            S.boxr[3 * i][t] = sk_rep(w[6 + i] & 0x1FF);
                            S.boxr[3 * i + 1][t] = sk_rep((w[6 + i] >> 10) & 0x1FF);
                            S.boxr[3 * i + 2][t] = sk_rep((w[6 + i] >> 20) & 0x1FF);
                        }
#pragma unroll
                        for (int rs = 0; rs < 27; rs++) {
                            const int cell0 = (rs / 3) * 9 + (rs % 3) * 3;       // compile-time
                            uint32_t m = 0;
#pragma unroll
                            for (int f = 0; f < 3; f++) {
                                const int c = cell0 + f;
                                const uint32_t bw = c < 32 ? b0 : (c < 64 ? b1 : b2);
                                if ((bw >> (c & 31)) & 1u) m |= 0x1FFu << (10 * f);
                            }
                            S.blk[rs][t] = m;
                        }
                        sp = 0; base_sp = 0;
                        p = sk_next_blank(b0, b1, b2, -1);
                        if (A.round != 0) {
                            // resume from the snapshot: re-assign the values chosen at levels 0..hi-1; the untried
                            // values of levels [lo, hi] are this piece's, everything shallower belongs to other pieces
                            const int lo = (int)(levels & 0xFF), hi = (int)((levels >> 8) & 0xFF);
                            const uint4* sb = A.snaps + (size_t)snap_id * kSnapWords;
                            uint4 cur = make_uint4(0, 0, 0, 0);
                            for (int l = 0; l <= hi; l++) {
                                if ((l & 7) == 0) cur = __ldg(sb + (l >> 3));
                                const uint32_t word = (l & 4) ? ((l & 2) ? cur.w : cur.z) : ((l & 2) ? cur.y : cur.x);
                                const uint32_t e = (l & 1) ? (word >> 16) : (word & 0xFFFF);
                                if (l == hi) { cand = e & 0x1FF; break; }
                                const uint32_t v = e >> 9;
                                const SkCell c = sk_decode(p);
                                const uint32_t bit = 1u << v;
                                S.rowr[c.r][t] |= sk_rep(bit);
                                S.boxr[c.box][t] |= sk_rep(bit);
                                S.colp[c.s][t] |= bit << (10 * c.f);
                                S.stk[l][t] = (uint16_t)(l >= lo ? e : (e & 0xFE00u));
                                p = sk_next_blank(b0, b1, b2, p);
                            }
                            sp = hi; base_sp = lo;
                        } else if (nblank == 0) {
                            // nothing to search: the givens are the solution (81 nodes)
                            A.nodes[puzzle] = 81; A.status[puzzle] = 1;
                            uint8_t* out = A.solution + (size_t)puzzle * A.stride;
                            const uint8_t* in = A.cells + (size_t)puzzle * A.stride;
                            for (int i = 0; i < 81; i++) out[i] = in[i];
                            have = false;
                        } else {
                            const SkCell c = sk_decode(p);
                            cand = ~(S.rowr[c.r][t] | (S.colp[c.s][t] >> (10 * c.f)) | S.boxr[c.box][t]) & 0x1FF;
                        }
                    }
                } else if (exhausted) done = true;
            }
            chunk_pos += min((unsigned long long)n_need, avail);
            if (__all_sync(0xFFFFFFFFu, done && !have)) break;
        }

        // ---------------- speculative tasks give up when an earlier subtree has the solution ----------------
        if (A.round != 0 && ((++poll & 255u) == 0)) {
            if (have && *(volatile unsigned long long*)(A.best_key + puzzle) < key_lo) {
                A.piece_nodes[task] = 0; A.piece_found[task] = 0; have = false;
            }
        }

        // ---------------- every value tried at this level: step back one level ----------------
        if (have && cand == 0) {
            if (sp == base_sp) {
                // the task's subtree is exhausted without a solution
                const unsigned long long tot = nodes_hi + nodes;
                if (A.round == 0) {
                    A.nodes[puzzle] = (81 - nblank) + tot;
                    A.status[puzzle] = 0;
                    uint8_t* out = A.solution + (size_t)puzzle * A.stride;
                    for (int i = 0; i < 81; i++) out[i] = 0;
                } else { A.piece_nodes[task] = tot; A.piece_found[task] = 0; }
                have = false;
            } else {
                --sp;
                const uint32_t e = S.stk[sp][t];
                cand = e & 0x1FF;
                const uint32_t bit = 1u << (e >> 9);
                p = sk_prev_blank(b0, b1, b2, p);
                const SkCell c = sk_decode(p);
                S.rowr[c.r][t] ^= sk_rep(bit);
                S.boxr[c.box][t] ^= sk_rep(bit);
                S.colp[c.s][t] ^= bit << (10 * c.f);
            }
        }

        // ---------------- AssignVar(next value) + forward check ----------------
        if (have && cand != 0) {
            const uint32_t bit = cand & (0u - cand);
            cand ^= bit;
            ++nodes;
            const SkCell c = sk_decode(p);
            const uint32_t rb = sk_rep(bit);
            const uint32_t rowv = S.rowr[c.r][t] | rb;
            uint32_t acc = 0;
            // later cells of the same row: own word (fields above f), then the stacks to the right
            {
                const uint32_t sel = (0xFFFFFFFFu << (10 * c.f + 10)) & SK_FULL3;
                const uint32_t u = rowv | S.colp[c.s][t] | S.boxr[c.box][t];
                acc |= (u & S.blk[c.r * 3 + c.s][t] & sel) + SK_ONES;
                for (int s2 = c.s + 1; s2 < 3; s2++) {
                    const uint32_t u2 = rowv | S.colp[s2][t] | S.boxr[c.band * 3 + s2][t];
                    acc |= (u2 & S.blk[c.r * 3 + s2][t]) + SK_ONES;
                }
            }
            // later rows: inside the band the whole box-row is a peer, below it only the column cell
            {
                const uint32_t colv = S.colp[c.s][t] | rb;
                const int in_band_last = c.band * 3 + 2;
                const uint32_t boxv = S.boxr[c.box][t] | rb;
                for (int r2 = c.r + 1; r2 <= in_band_last; r2++) {
                    const uint32_t u = S.rowr[r2][t] | colv | boxv;
                    acc |= (u & S.blk[r2 * 3 + c.s][t]) + SK_ONES;
                }
                const uint32_t fsel = 0x1FFu << (10 * c.f);
                for (int r2 = in_band_last + 1; r2 < 9; r2++) {
                    const int box2 = ((r2 * 11) >> 5) * 3 + c.s;
                    const uint32_t u = S.rowr[r2][t] | colv | S.boxr[box2][t];
                    acc |= (u & S.blk[r2 * 3 + c.s][t] & fsel) + SK_ONES;
                }
            }
            if ((acc & SK_SPARE) == 0) {
                // no wipe-out: the value stands
                const uint32_t v = __ffs(bit) - 1;
                if (sp == nblank - 1) {
                    // last blank assigned: the DFS-first solution of this subtree
                    S.stk[sp][t] = (uint16_t)(v << 9);
                    const unsigned long long tot = nodes_hi + nodes;
                    uint8_t* out;
                    if (A.round == 0) {
                        const unsigned long long all = (81 - nblank) + tot;
                        const bool over = A.user_budget && all > A.user_budget;
                        A.nodes[puzzle] = over ? A.user_budget + 1 : all;
                        A.status[puzzle] = over ? 2 : 1;
                        out = A.solution + (size_t)puzzle * A.stride;
                        if (over) { for (int i = 0; i < 81; i++) out[i] = 0; out = nullptr; }
                    } else {
                        A.piece_nodes[task] = tot; A.piece_found[task] = 1;
                        atomicMin(A.best_key + puzzle, key_lo);
                        out = A.piece_sol + (size_t)task * 81;
                    }
                    if (out) {
                        const uint8_t* in = A.cells + (size_t)puzzle * A.stride;
                        int l = 0;
                        for (int i = 0; i < 81; i++) {
                            const uint32_t bw = i < 32 ? b0 : (i < 64 ? b1 : b2);
                            if ((bw >> (i & 31)) & 1u) { out[i] = (uint8_t)((S.stk[l][t] >> 9) + 1); ++l; }
                            else out[i] = in[i];
                        }
                    }
                    have = false;
                } else {
                    S.rowr[c.r][t] = rowv;
                    S.boxr[c.box][t] |= rb;
                    S.colp[c.s][t] |= bit << (10 * c.f);
                    S.stk[sp][t] = (uint16_t)(cand | (v << 9));
                    ++sp;
                    p = sk_next_blank(b0, b1, b2, p);
                    const SkCell n = sk_decode(p);
                    cand = ~(S.rowr[n.r][t] | (S.colp[n.s][t] >> (10 * n.f)) | S.boxr[n.box][t]) & 0x1FF;
                }
            }
            // ---------------- node budget of the task ----------------
            if (nodes >= 0x80000000u) { nodes_hi += nodes; nodes = 0; }      // only unlimited tasks get here
            if (have && nodes >= limit) {
                if (stop_is_budget) {
                    A.nodes[puzzle] = A.user_budget + 1; A.status[puzzle] = 2;
                    uint8_t* out = A.solution + (size_t)puzzle * A.stride;
                    for (int i = 0; i < 81; i++) out[i] = 0;
                    have = false;
                } else {
                    // Split: the untried values left on the stack, levels [base_sp, sp], are cut into up to four
                    // contiguous level groups.  DFS order is deepest first, so the deepest group continues this
                    // task's own search and gets the lowest key interval; the shallowest levels — the largest
                    // subtrees — get a group each.
                    int m = cand ? 1 : 0;
                    for (int l = base_sp; l < sp; l++) m += (S.stk[l][t] & 0x1FF) ? 1 : 0;
                    const int groups = m < 4 ? m : 4;
                    const unsigned long long width = (key_hi - key_lo) / (unsigned long long)(groups + 1);
                    unsigned long long slot = 0, sslot = 0;
                    bool ok = m > 0 && width > 0;
                    if (ok) {
                        slot = atomicAdd(A.ctrl + 1, (unsigned long long)groups);
                        sslot = atomicAdd(A.ctrl + 2, 1ull);
                        ok = slot + groups <= A.piece_cap && sslot < A.snap_cap;
                        if (!ok)                                                 // pool full: fill what was reserved with null pieces
                            for (unsigned long long i = slot; i < slot + groups && i < A.piece_cap; i++) {
                                uint4* pr = reinterpret_cast<uint4*>(A.pieces + i);
                                pr[0] = make_uint4(puzzle, 0u, 0xFFFFFFFFu, 0u);
                                pr[1] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                            }
                    }
                    if (!ok) limit = 0xFFFFFFFFu;          // nothing to hand over, or key space / pool exhausted: finish unsplit
                    else {
                        // stack snapshot: levels 0..sp, entry sp = the current cell's untried values
                        uint4* sb = A.snaps + (size_t)sslot * kSnapWords;
                        for (int q = 0; q * 8 <= sp; q++) {
                            uint32_t w4[4];
#pragma unroll
                            for (int i = 0; i < 4; i++) {
                                const int l0 = q * 8 + 2 * i, l1 = l0 + 1;
                                const uint32_t e0 = l0 < sp ? (uint32_t)S.stk[l0][t] : (l0 == sp ? cand : 0u);
                                const uint32_t e1 = l1 < sp ? (uint32_t)S.stk[l1][t] : (l1 == sp ? cand : 0u);
                                w4[i] = e0 | (e1 << 16);
                            }
                            sb[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                        // group g (0 = shallowest) takes 1, 1, 2 non-empty levels; the deepest group takes the rest
                        unsigned long long k = key_lo + width * (unsigned long long)groups;   // shallowest = last in DFS order
                        int l = base_sp;
                        int left = m;
                        for (int g = 0; g < groups; g++) {
                            int take = g == groups - 1 ? left : (g == 2 ? 2 : 1);
                            if (take > left - (groups - 1 - g)) take = left - (groups - 1 - g);
                            int lo_l = -1, hi_l = -1;
                            while (take > 0) {
                                const uint32_t cl = l == sp ? cand : (uint32_t)(S.stk[l][t] & 0x1FF);
                                if (cl) { if (lo_l < 0) lo_l = l; hi_l = l; --take; --left; }
                                ++l;
                            }
                            uint4* pr = reinterpret_cast<uint4*>(A.pieces + slot + g);
                            pr[0] = make_uint4(puzzle, (uint32_t)sslot, (uint32_t)lo_l | ((uint32_t)hi_l << 8), 0u);
                            pr[1] = make_uint4((uint32_t)k, (uint32_t)(k >> 32), (uint32_t)(k + width), (uint32_t)((k + width) >> 32));
                            k -= width;
                        }
                        const unsigned long long tot = nodes_hi + nodes;
                        if (A.round == 0) {
                            A.nodes[puzzle] = (81 - nblank) + tot;          // the part before every piece; pieces add to it
                            A.status[puzzle] = SK_STATUS_SPLIT;
                        } else { A.piece_nodes[task] = tot; A.piece_found[task] = 0; }
                        have = false;
                    }
                }
            }
        }
    }
}

// After the last round.  Pass 1: every piece that holds a solution has already lowered best_key
// (atomicMin in the search).  Pass 2, one thread per piece: add its nodes if it lies at or before
// the puzzle's first solution in DFS order; the piece AT the first solution writes it out.
__global__ void k_sudoku_account(SudokuArgs A, unsigned long long n_pieces) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pieces) return;
    const SudokuPiece pc = A.pieces[i];
    if (pc.levels == 0xFFFFFFFFu) return;
    const unsigned long long best = A.best_key[pc.puzzle];
    if (pc.key_lo > best) return;
    atomicAdd(A.nodes + pc.puzzle, A.piece_nodes[i]);
    if (pc.key_lo == best && A.piece_found[i]) {
        uint8_t* out = A.solution + (size_t)pc.puzzle * A.stride;
        const uint8_t* in = A.piece_sol + (size_t)i * 81;
        for (int c = 0; c < 81; c++) out[c] = in[c];
        A.status[pc.puzzle] = 1;
    }
}

// Pass 3, one thread per instance: split puzzles without a winner are UNSAT; the API node budget is
// applied to the exact totals; batch totals.
__global__ void k_sudoku_finish(SudokuArgs A, unsigned long long* totals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned sat = 0, unsat = 0, budget = 0;
    unsigned long long nd = 0;
    if (i < A.n) {
        uint8_t st = A.status[i];
        if (st != SK_STATUS_DEFER) {
            if (st == SK_STATUS_SPLIT) {
                st = 0;
                uint8_t* out = A.solution + (size_t)i * A.stride;
                for (int c = 0; c < 81; c++) out[c] = 0;
            }
            if (A.user_budget && A.nodes[i] > A.user_budget && st != 2) {
                st = 2;
                A.nodes[i] = A.user_budget + 1;
                uint8_t* out = A.solution + (size_t)i * A.stride;
                for (int c = 0; c < 81; c++) out[c] = 0;
            }
            A.status[i] = st;
            nd = A.nodes[i];
            sat = st == 1; unsat = st == 0; budget = st == 2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sat += __shfl_down_sync(0xFFFFFFFFu, sat, o);
        unsat += __shfl_down_sync(0xFFFFFFFFu, unsat, o);
        budget += __shfl_down_sync(0xFFFFFFFFu, budget, o);
        nd += __shfl_down_sync(0xFFFFFFFFu, nd, o);
    }
    if ((threadIdx.x & 31) == 0 && (sat | unsat | budget | (nd != 0))) {
        if (sat) atomicAdd(totals + 0, (unsigned long long)sat);
        if (unsat) atomicAdd(totals + 1, (unsigned long long)unsat);
        if (budget) atomicAdd(totals + 2, (unsigned long long)budget);
        atomicAdd(totals + 3, nd);
    }
}

// Compacts the ids of the instances the digest deferred to the generic engine.
__global__ void k_sudoku_collect_deferred(const uint8_t* __restrict__ status, long long n, int* __restrict__ list,
                                          unsigned long long* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == SK_STATUS_DEFER) list[atomicAdd(count, 1ull)] = (int)i;
}

}  // namespace dq
