// tests/cpp/dropin_scenarios.cpp — the same model source compiled twice:
//   * against the UNMODIFIED reference header (-I/root/reference) in the build container, to
//     produce tests/golden/dropin_reference.json (tests/golden/make_dropin_golden.sh);
//   * against include/dequan.h + libdequan_b200.so on the GPU box (tests/test_dropin_cpp.py).
// It uses only the public API both headers share (reference dequan.h:52-355) and prints one JSON
// object per scenario: outcome, solution, stats.assigned_vars, assign_order, and the complete
// post-solve Assignment state (current_domains, saved_domains).
// The first three scenarios are the reference's own test program (test/main-test.cpp:27-233).
#include <climits>
#include <new>
#include <utility>
#include <cstdio>
#include <string>

#define DEQUAN_USE_STDVECTOR
#define DEQUAN_WITH_STATS
#define DEQUAN_IMPLEMENTATION
#include "dequan.h"

using namespace dequan;

// A user-defined, check-only binary constraint (only LinkVars + Evaluate): |x - y| != gap.
struct GapConstraint : public Constraint {
    GapConstraint(VarId a, VarId b, int g) : x(a), y(b), gap(g) {}
    virtual void LinkVars(Array<Var>& vars) {
        vars[x].linked_constraints.push_back(this);
        vars[y].linked_constraints.push_back(this);
    }
    virtual Eval Evaluate(const Array<InstVar>& iv, VarId) {
        if (iv[x].value == InstVar::UNASSIGNED || iv[y].value == InstVar::UNASSIGNED) return Eval::NA;
        int d = iv[x].value - iv[y].value;
        if (d < 0) d = -d;
        return d != gap ? Eval::Passed : Eval::Failed;
    }
    VarId x, y;
    int gap;
};

// A user-defined constraint WITH its own filtering step (AplyArcConsistency, reference dequan.h:145-147): x + gap <= y.
struct OrderedConstraint : public Constraint {
    OrderedConstraint(VarId a, VarId b, int g) : x(a), y(b), gap(g) {}
    virtual void LinkVars(Array<Var>& vars) {
        vars[x].linked_constraints.push_back(this);
        vars[y].linked_constraints.push_back(this);
    }
    virtual Eval Evaluate(const Array<InstVar>& iv, VarId) {
        if (iv[x].value == InstVar::UNASSIGNED || iv[y].value == InstVar::UNASSIGNED) return Eval::NA;
        return iv[x].value + gap <= iv[y].value ? Eval::Passed : Eval::Failed;
    }
    virtual bool AplyArcConsistency(Assignment& a, VarId) {
        const int xv = a.inst_vars[x].value, yv = a.inst_vars[y].value;
        if (xv != InstVar::UNASSIGNED && yv == InstVar::UNASSIGNED) {
            Domain& d = a.current_domains[y];
            a.EnsureSavedDomain(y, d);
            d.ExcludeInf(xv + gap);
            return d.Size() > 0;
        }
        if (yv != InstVar::UNASSIGNED && xv == InstVar::UNASSIGNED) {
            Domain& d = a.current_domains[x];
            a.EnsureSavedDomain(x, d);
            d.ExcludeSup(yv - gap + 1);
            return d.Size() > 0;
        }
        return true;
    }
    VarId x, y;
    int gap;
};

static void print_ints(const Array<int>& v) {
    printf("[");
    for (size_t i = 0; i < v.size(); i++) printf("%s%d", i ? "," : "", v[i]);
    printf("]");
}

static void report(const char* name, bool ok, const CSP& csp, const Assignment& a, bool dump_state = true) {
    printf("{\"name\":\"%s\",\"ok\":%s,\"assigned_vars\":%llu,\"assigned_var_count\":%d,\"values\":[", name, ok ? "true" : "false",
           a.stats.assigned_vars, a.assigned_var_count);
    for (size_t v = 0; v < csp.vars.size(); v++) printf("%s%d", v ? "," : "", a.GetInstVarValue((VarId)v));
    printf("],\"assign_order\":");
    print_ints(a.assign_order);
    if (dump_state) {
        printf(",\"current_domains\":[");
        for (size_t v = 0; v < a.current_domains.size(); v++) {
            printf("%s{\"t\":%d,\"v\":", v ? "," : "", (int)a.current_domains[v].type);
            print_ints(a.current_domains[v].values);
            printf("}");
        }
        printf("],\"saved_domains\":[");
        for (size_t d = 0; d < a.saved_domains.size(); d++) {
            printf("%s[", d ? "," : "");
            const Array<SavedDomain>& f = a.saved_domains[d].domains;
            for (size_t i = 0; i < f.size(); i++) {
                printf("%s{\"id\":%d,\"t\":%d,\"v\":", i ? "," : "", f[i].var_id, (int)f[i].type);
                print_ints(f[i].values);
                printf("}");
            }
            printf("]");
        }
        printf("]");
    }
    printf("}\n");
}

static void queens_model(CSP& csp, int n) {
    Array<VarId> q(n);
    for (int i = 0; i < n; i++) q[i] = csp.AddIntVar(0, n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, 0));
            csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, j - i));
            csp.AddConstraint(OpConstraint(q[i], q[j], OpConstraint::Op::NotEqual, i - j));
        }
    csp.FinalizeModel();
}

static void queens(const char* name, int n, bool reverse_order = false) {
    CSP csp;
    queens_model(csp, n);
    Assignment a;
    a.Reset(csp);
    if (reverse_order)   // assign_order is a public field: a caller may edit it between Reset and the solve
        for (int i = 0; i < n / 2; i++) std::swap(a.assign_order[i], a.assign_order[n - 1 - i]);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
    if (ok) {            // a second call on a complete assignment answers true without searching
        bool again = csp.ForwardCheckingStep(a);
        std::string nm = std::string(name) + "_again";
        report(nm.c_str(), again, csp, a, false);
    }
}

static const int kGrid[81] = {   // test/main-test.cpp:92-105
    0, 0, 3, 0, 2, 0, 6, 0, 0, 9, 0, 0, 3, 0, 5, 0, 0, 1, 0, 0, 1, 8, 0, 6, 4, 0, 0,
    0, 0, 8, 1, 0, 2, 9, 0, 0, 7, 0, 0, 0, 0, 0, 0, 0, 8, 0, 0, 6, 7, 0, 8, 2, 0, 0,
    0, 0, 2, 6, 0, 9, 5, 0, 0, 8, 0, 0, 2, 0, 3, 0, 0, 9, 0, 0, 5, 0, 1, 0, 3, 0, 0};

static void sudoku(const char* name, bool boxes, bool binary, int break_cell = -1) {
    CSP csp;
    Array<VarId> cell(81);
    for (int i = 0; i < 81; i++) {
        int g = kGrid[i];
        if (i == break_cell) g = 3;          // a given that contradicts its row -> no solution
        cell[i] = g ? csp.AddFixedVar(g) : csp.AddIntVar(1, 10);
    }
    Array<Array<VarId> > groups;
    for (int r = 0; r < 9; r++) { Array<VarId> g; for (int c = 0; c < 9; c++) g.push_back(cell[r * 9 + c]); groups.push_back(g); }
    for (int c = 0; c < 9; c++) { Array<VarId> g; for (int r = 0; r < 9; r++) g.push_back(cell[r * 9 + c]); groups.push_back(g); }
    if (boxes)
        for (int b = 0; b < 9; b++) {
            Array<VarId> g;
            for (int k = 0; k < 9; k++) g.push_back(cell[(b / 3 * 3 + k / 3) * 9 + (b % 3 * 3 + k % 3)]);
            groups.push_back(g);
        }
    for (size_t gi = 0; gi < groups.size(); gi++) {
        if (!binary) csp.AddConstraint(AllDifferentConstraint(groups[gi]));
        else
            for (int i = 0; i < 9; i++)
                for (int j = i + 1; j < 9; j++) csp.AddConstraint(OpConstraint(groups[gi][i], groups[gi][j], OpConstraint::Op::NotEqual, 0));
    }
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

static void op_inequality() {   // test/main-test.cpp:187-233
    CSP csp;
    VarId v0 = csp.AddIntVar(0, 10), v1 = csp.AddIntVar(0, 10), v2 = csp.AddFixedVar(6), v3 = csp.AddFixedVar(5);
    csp.AddConstraint(OpConstraint(v0, v2, OpConstraint::Op::Inf, 0));
    csp.AddConstraint(OpConstraint(v0, v3, OpConstraint::Op::SupEqual, 0));
    csp.AddConstraint(OpConstraint(v1, v2, OpConstraint::Op::InfEqual, 0));
    csp.AddConstraint(OpConstraint(v1, v3, OpConstraint::Op::Sup, 0));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report("op_inequality", ok, csp, a);
}

// Values domains in non-ascending order, negative values, offsets, Equal / EqualityConstraint
// (weak Intersect), OrRange (check-only) and an all-different on top.
static void mixed_ops(const char* name, int r3_gap_lo) {
    CSP csp;
    int d0[] = {4, -2, 7, 1, 0}, d1[] = {3, 9, -1, 5}, d2[] = {6, 2, 8};
    VarId a0 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d0, d0 + 5)));
    VarId a1 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d1, d1 + 4)));
    VarId a2 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d2, d2 + 3)));
    int r3[] = {-3, r3_gap_lo, 2, 6};
    VarId a3 = csp.AddIntVar(Domain(DomainType::Ranges, Array<int>(r3, r3 + 4)));
    VarId a4 = csp.AddIntVar(-2, 5);
    VarId a5 = csp.AddBoolVar();
    VarId a6 = csp.AddIntVar(0, 12);
    csp.AddConstraint(OpConstraint(a0, a1, OpConstraint::Op::Inf, -2));        // a0 < a1 - 2
    csp.AddConstraint(OpConstraint(a2, a0, OpConstraint::Op::SupEqual, 5));    // a2 >= a0 + 5
    csp.AddConstraint(OpConstraint(a3, a4, OpConstraint::Op::Equal, -1));      // a3 == a4 - 1
    csp.AddConstraint(EqualityConstraint(a5, a3));                             // a5 == a3
    csp.AddConstraint(OpConstraint(a6, a2, OpConstraint::Op::Sup, 3));         // a6 > a2 + 3
    csp.AddConstraint(OpConstraint(a6, a1, OpConstraint::Op::NotEqual, 2));    // a6 != a1 + 2
    csp.AddConstraint(OrRangeConstraint(a4, a6, 10, 12));                      // a4 in [10,12) or a6 in [10,12)
    Array<VarId> ad;
    ad.push_back(a0); ad.push_back(a3); ad.push_back(a4);
    csp.AddConstraint(AllDifferentConstraint(ad));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

// Domains of more than 32 values: six queens on a 48-column board with an ordering chain on top, and a 60-value
// Values domain listed in descending order.
static void wide_domains(const char* name, int shift, bool contradict) {
    CSP csp;
    const int n = 6, cols = 48;
    for (int i = 0; i < n; i++) csp.AddIntVar(0, cols);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, 0));
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, j - i));
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, i - j));
        }
    Array<int> desc;
    for (int v = 59; v >= 0; v--) desc.push_back(v + shift);
    VarId w = csp.AddIntVar(Domain(DomainType::Values, desc));
    csp.AddConstraint(OpConstraint(w, 0, OpConstraint::Op::Sup, 40));           // w > q0 + 40
    csp.AddConstraint(OpConstraint(5, w, OpConstraint::Op::Equal, -30));        // q5 == w - 30
    csp.AddConstraint(OpConstraint(2, 3, OpConstraint::Op::InfEqual, -20));     // q2 <= q3 - 20
    if (contradict) csp.AddConstraint(OpConstraint(1, 0, OpConstraint::Op::Equal, 0));   // q1 == q0 against q0 != q1
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

// 7 variables on [0,7), pairwise different, plus the user-defined gap constraint between neighbours.
static void user_constraint(const char* name, int gap) {
    CSP csp;
    const int n = 7;
    Array<VarId> v(n);
    for (int i = 0; i < n; i++) v[i] = csp.AddIntVar(0, n);
    csp.AddConstraint(AllDifferentConstraint(v));
    for (int i = 0; i + 1 < n; i++) csp.AddConstraint(GapConstraint(v[i], v[i + 1], gap));
    for (int i = 0; i + 2 < n; i++) csp.AddConstraint(GapConstraint(v[i], v[i + 2], gap + 1));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

// A partially assigned Assignment: the caller assigns the first variables of assign_order by hand (AssignVar, no
// filtering) and the solve resumes at assign_order[assigned_var_count] (reference dequan.h:411-414, 504).
static void resume(const char* name, CSP& csp, const int* vals, int k) {
    Assignment a;
    a.Reset(csp);
    for (int i = 0; i < k; i++) a.AssignVar(a.assign_order[i], vals[i]);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

static void queens_resume(const char* name, int n, const int* vals, int k) {
    CSP csp;
    queens_model(csp, n);
    resume(name, csp, vals, k);
}

static void mixed_resume(const char* name, const int* vals, int k) {
    CSP csp;
    int d0[] = {4, -2, 7, 1, 0}, d1[] = {3, 9, -1, 5}, d2[] = {6, 2, 8};
    VarId a0 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d0, d0 + 5)));
    VarId a1 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d1, d1 + 4)));
    VarId a2 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d2, d2 + 3)));
    VarId a3 = csp.AddIntVar(-3, 6);
    VarId a4 = csp.AddIntVar(-2, 5);
    VarId a5 = csp.AddBoolVar();
    csp.AddConstraint(OpConstraint(a0, a1, OpConstraint::Op::Inf, -2));
    csp.AddConstraint(OpConstraint(a2, a0, OpConstraint::Op::SupEqual, 5));
    csp.AddConstraint(OpConstraint(a3, a4, OpConstraint::Op::Equal, -1));
    csp.AddConstraint(EqualityConstraint(a5, a3));
    csp.AddConstraint(OrRangeConstraint(a4, a2, 6, 9));
    csp.AddConstraint(GapConstraint(a1, a2, 1));
    Array<VarId> ad;
    ad.push_back(a0); ad.push_back(a3); ad.push_back(a4); ad.push_back(a5);
    csp.AddConstraint(AllDifferentConstraint(ad));
    csp.FinalizeModel();
    resume(name, csp, vals, k);
}

static void sudoku_resume(const char* name, int k) {     // the first k givens assigned by hand instead of by the search
    CSP csp;
    Array<VarId> cell(81);
    for (int i = 0; i < 81; i++) cell[i] = kGrid[i] ? csp.AddFixedVar(kGrid[i]) : csp.AddIntVar(1, 10);
    for (int r = 0; r < 9; r++) { Array<VarId> g; for (int c = 0; c < 9; c++) g.push_back(cell[r * 9 + c]); csp.AddConstraint(AllDifferentConstraint(g)); }
    for (int c = 0; c < 9; c++) { Array<VarId> g; for (int r = 0; r < 9; r++) g.push_back(cell[r * 9 + c]); csp.AddConstraint(AllDifferentConstraint(g)); }
    for (int b = 0; b < 9; b++) {
        Array<VarId> g;
        for (int j = 0; j < 9; j++) g.push_back(cell[(b / 3 * 3 + j / 3) * 9 + (b % 3 * 3 + j % 3)]);
        csp.AddConstraint(AllDifferentConstraint(g));
    }
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    for (int i = 0; i < k; i++) a.AssignVar(a.assign_order[i], csp.domains[a.assign_order[i]].values[0]);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

// The array macros of the reference's std::vector mode (dequan.h:32-42) in user code.
static void array_macros() {
    Array<int> vals;
    DEQUAN_Array_PushBack(vals, 5); DEQUAN_Array_PushBack(vals, 3); DEQUAN_Array_PushBack(vals, 9); DEQUAN_Array_PushBack(vals, 7);
    DEQUAN_Array_Sort(vals, [](int x, int y) { return x < y; });     // 3 5 7 9
    DEQUAN_Array_Insert(vals, 1, 4);                                  // 3 4 5 7 9
    DEQUAN_Array_Erase(vals, 3, 4)                                    // 3 4 5 9 (the macro brings its own semicolon)
    CSP csp;
    VarId x = csp.AddIntVar(Domain(DomainType::Values, vals));
    VarId y = csp.AddIntVar(Domain(DomainType::Values, vals));
    VarId z = csp.AddIntVar(Domain(DomainType::Values, vals));
    csp.AddConstraint(OpConstraint(x, y, OpConstraint::Op::Sup, 1));
    csp.AddConstraint(OpConstraint(z, x, OpConstraint::Op::Sup, 3));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report("array_macros", ok && DEQUAN_Array_Size(vals) == 4 && DEQUAN_Array_Back(vals) == 9, csp, a);
}

// Filtering user constraints next to built-in ones: a chain v0 + g <= v1 + ... with an all-different on top and a
// check-only gap constraint; `reverse_domains` lists the Values domains in descending order (iteration order matters).
static void user_filter(const char* name, int n, int g, bool reverse_domains, int extra_gap) {
    CSP csp;
    Array<VarId> v(n);
    for (int i = 0; i < n; i++) {
        if (i % 2 == 0) v[i] = csp.AddIntVar(0, n + 3);
        else {
            Array<int> vals;
            for (int k = 0; k < n + 3; k++) vals.push_back(reverse_domains ? n + 2 - k : k);
            v[i] = csp.AddIntVar(Domain(DomainType::Values, vals));
        }
    }
    for (int i = 0; i + 1 < n; i++) csp.AddConstraint(OrderedConstraint(v[i], v[i + 1], (i % 2) ? g : 1));
    csp.AddConstraint(AllDifferentConstraint(v));
    if (extra_gap) csp.AddConstraint(GapConstraint(v[0], v[n - 1], extra_gap));
    csp.AddConstraint(OpConstraint(v[1], v[n - 2], OpConstraint::Op::NotEqual, -3));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

// Values domains that list a value more than once (every copy is visited; Exclude erases the first copy only).
static void duplicate_values(const char* name, bool with_equal) {
    CSP csp;
    int d0[] = {3, 1, 3, 2, 1}, d1[] = {2, 2, 3, 1}, d2[] = {1, 3, 3, 3, 2, 2};
    VarId a0 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d0, d0 + 5)));
    VarId a1 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d1, d1 + 4)));
    VarId a2 = csp.AddIntVar(Domain(DomainType::Values, Array<int>(d2, d2 + 6)));
    VarId a3 = csp.AddIntVar(1, 4);
    csp.AddConstraint(OpConstraint(a0, a1, OpConstraint::Op::NotEqual, 0));
    csp.AddConstraint(OpConstraint(a1, a2, OpConstraint::Op::NotEqual, 1));
    csp.AddConstraint(OpConstraint(a0, a2, OpConstraint::Op::NotEqual, 0));
    csp.AddConstraint(OpConstraint(a0, a2, OpConstraint::Op::NotEqual, 0));      // the same exclusion twice: two copies go
    Array<VarId> ad;
    ad.push_back(a1); ad.push_back(a2); ad.push_back(a3);
    csp.AddConstraint(AllDifferentConstraint(ad));
    if (with_equal) csp.AddConstraint(EqualityConstraint(a3, a0));
    csp.AddConstraint(OpConstraint(a3, a2, OpConstraint::Op::Sup, 0));
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report(name, ok, csp, a);
}

static void empty_model() {
    CSP csp;
    csp.FinalizeModel();
    Assignment a;
    a.Reset(csp);
    bool ok = csp.ForwardCheckingStep(a);
    report("empty_model", ok, csp, a);
}

// Host-side Domain operations (public API, reference dequan.h:941-1172): a scripted pseudo-random
// sequence over Values and Ranges domains.  Needs no device.
static void domain_ops() {
    unsigned long long rng = 88172645463325252ull;
    auto next = [&rng](int mod) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (int)(rng % (unsigned)mod); };
    for (int trial = 0; trial < 400; trial++) {
        Domain d;
        if (trial & 1) {
            d.type = DomainType::Ranges;
            int lo = next(6) - 3;
            const int nr = 1 + next(3);
            for (int r = 0; r < nr; r++) { int len = 1 + next(5); d.values.push_back(lo); d.values.push_back(lo + len); lo += len + 1 + next(3); }
        } else {
            d.type = DomainType::Values;
            const int n = 1 + next(8);
            for (int i = 0; i < n; i++) {
                int v = next(14) - 3;
                bool dup = false;
                for (size_t j = 0; j < d.values.size(); j++) dup |= d.values[j] == v;
                if (!dup) d.values.push_back(v);
            }
        }
        printf("{\"name\":\"domain_ops_%d\",\"steps\":[", trial);
        for (int step = 0; step < 6; step++) {
            const int op = next(7), x = next(16) - 4, y = next(16) - 4;
            switch (op) {
                case 0: d.Intersect(x); break;
                case 1: d.Exclude(x); break;
                case 2: d.ExcludeSup(x); break;
                case 3: d.ExcludeInf(x); break;
                case 4: d.IntersectRange(x < y ? x : y, x < y ? y : x); break;
                case 5: if (x != y) d.Intersect(x < y ? x : y, x < y ? y : x); break;
                default: d.Exclude(y); break;
            }
            printf("%s{\"op\":%d,\"x\":%d,\"y\":%d,\"t\":%d,\"size\":%d,\"v\":", step ? "," : "", op, x, y, (int)d.type, d.Size());
            print_ints(d.values);
            printf("}");
        }
        printf("]}\n");
    }
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "domains") { domain_ops(); return 0; }
    op_inequality();
    queens("queens8", 8);
    sudoku("sudoku_rows_cols_alldiff", false, false);     // the reference's SudokuTest (no boxes)
    queens("queens3_unsat", 3);
    queens("queens6", 6);
    queens("queens12", 12);
    queens("queens9_reversed_order", 9, true);
    sudoku("sudoku_boxes_alldiff", true, false);
    sudoku("sudoku_boxes_binary", true, true);
    sudoku("sudoku_contradictory_given", true, false, 0);
    mixed_ops("mixed_ops_unsat", 0);
    mixed_ops("mixed_ops_sat", 1);
    user_constraint("user_gap1", 1);
    user_constraint("user_gap2", 2);
    empty_model();
    wide_domains("wide_domains_sat", 10, false);
    wide_domains("wide_domains_shift40", 40, false);
    wide_domains("wide_domains_unsat", 10, true);
    { const int v[] = {3}; queens_resume("resume_queens8_first3", 8, v, 1); }
    { const int v[] = {0, 2, 4}; queens_resume("resume_queens8_prefix024", 8, v, 3); }
    { const int v[] = {0, 1}; queens_resume("resume_queens8_conflicting_prefix", 8, v, 2); }   // never checked against each other
    { const int v[] = {5, 5}; queens_resume("resume_queens6_same_row", 6, v, 2); }
    { const int v[] = {100}; queens_resume("resume_queens6_foreign_value", 6, v, 1); }
    { const int v[] = {1}; queens_resume("resume_queens3_unsat", 3, v, 1); }
    { const int v[] = {0, 1, 2, 3, 4, 5}; queens_resume("resume_queens6_all_assigned", 6, v, 6); }
    { const int v[] = {1, 8}; mixed_resume("resume_mixed_two", v, 2); }
    { const int v[] = {0, 6, 9}; mixed_resume("resume_mixed_three", v, 3); }
    { const int v[] = {1, 2, 3, 0}; mixed_resume("resume_mixed_unsat", v, 4); }
    sudoku_resume("resume_sudoku_10_givens", 10);
    sudoku_resume("resume_sudoku_all_givens", 32);
    array_macros();
    user_filter("user_filter_sat", 5, 2, false, 0);
    user_filter("user_filter_reversed_domains", 6, 1, true, 4);
    user_filter("user_filter_unsat", 6, 3, false, 0);
    user_filter("user_filter_gap", 5, 1, true, 6);
    duplicate_values("duplicate_values", false);
    duplicate_values("duplicate_values_equal", true);
    return 0;
}
