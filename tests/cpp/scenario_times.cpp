// tests/cpp/scenario_times.cpp — wall time of the reference's own three scenarios (test/main-test.cpp:27-233) through
// whichever dequan.h is on the include path: the unmodified reference (CPU recursion) or include/dequan.h (B200).
// Prints one JSON line per scenario: first call (for the drop-in: model lowering + dq_compile + table upload + solve),
// and the best of five repeats on a fresh Assignment (the drop-in re-uses the compiled model cached in the CSP).
#include <climits>
#include <new>
#include <utility>
#include <chrono>
#include <cstdio>

#define DEQUAN_USE_STDVECTOR
#define DEQUAN_WITH_STATS
#define DEQUAN_IMPLEMENTATION
#include "dequan.h"

using namespace dequan;

static double solve_ms(CSP& csp, unsigned long long* nodes, bool* ok) {
    Assignment a;
    a.Reset(csp);
    const auto t0 = std::chrono::steady_clock::now();
    *ok = csp.ForwardCheckingStep(a);
    const auto t1 = std::chrono::steady_clock::now();
    *nodes = a.stats.assigned_vars;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

static void timed(const char* name, CSP& csp) {
    unsigned long long nodes = 0;
    bool ok = false;
    const double first = solve_ms(csp, &nodes, &ok);
    double best = 1e30;
    for (int i = 0; i < 5; i++) { const double t = solve_ms(csp, &nodes, &ok); if (t < best) best = t; }
    printf("{\"scenario\":\"%s\",\"ok\":%s,\"nodes\":%llu,\"first_call_ms\":%.4f,\"repeat_ms\":%.4f}\n", name, ok ? "true" : "false", nodes, first, best);
}

int main() {
    {   // OpInequalityTest, main-test.cpp:187-233
        CSP csp;
        VarId v0 = csp.AddIntVar(0, 10), v1 = csp.AddIntVar(0, 10), c6 = csp.AddFixedVar(6), c5 = csp.AddFixedVar(5);
        csp.AddConstraint(OpConstraint(v0, c6, OpConstraint::Op::Inf, 0));
        csp.AddConstraint(OpConstraint(v0, c5, OpConstraint::Op::SupEqual, 0));
        csp.AddConstraint(OpConstraint(v1, c6, OpConstraint::Op::InfEqual, 0));
        csp.AddConstraint(OpConstraint(v1, c5, OpConstraint::Op::Sup, 0));
        csp.FinalizeModel();
        timed("OpInequalityTest", csp);
    }
    {   // NQueensTest(8), main-test.cpp:27-86
        CSP csp;
        const int n = 8;
        Array<VarId> q(n);
        for (int i = 0; i < n; i++) q[i] = csp.AddIntVar(0, n);
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, 0));
                csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, j - i));
                csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, i - j));
            }
        csp.FinalizeModel();
        timed("NQueensTest8", csp);
    }
    {   // SudokuTest, main-test.cpp:88-185: rows and columns only (no boxes), AllDifferent
        static const int grid[81] = {
            0, 0, 3, 0, 2, 0, 6, 0, 0, 9, 0, 0, 3, 0, 5, 0, 0, 1, 0, 0, 1, 8, 0, 6, 4, 0, 0,
            0, 0, 8, 1, 0, 2, 9, 0, 0, 7, 0, 0, 0, 0, 0, 0, 0, 8, 0, 0, 6, 7, 0, 8, 2, 0, 0,
            0, 0, 2, 6, 0, 9, 5, 0, 0, 8, 0, 0, 2, 0, 3, 0, 0, 9, 0, 0, 5, 0, 1, 0, 3, 0, 0};
        CSP csp;
        Array<VarId> cell(81);
        for (int i = 0; i < 81; i++) cell[i] = grid[i] ? csp.AddFixedVar(grid[i]) : csp.AddIntVar(1, 10);
        for (int r = 0; r < 9; r++) { Array<VarId> g; for (int c = 0; c < 9; c++) g.push_back(cell[r * 9 + c]); csp.AddConstraint(AllDifferentConstraint(g)); }
        for (int c = 0; c < 9; c++) { Array<VarId> g; for (int r = 0; r < 9; r++) g.push_back(cell[r * 9 + c]); csp.AddConstraint(AllDifferentConstraint(g)); }
        csp.FinalizeModel();
        timed("SudokuTest_rows_cols", csp);
    }
    return 0;
}
