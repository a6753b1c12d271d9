// The dequan::b200 extensions of the drop-in header (no counterpart in the reference): CountAll, EnumerateAll, Solve.
// Prints one JSON line per model; tests/test_dropin_cpp.py checks them against the oracle and the enumeration goldens.
#define DEQUAN_USE_STDVECTOR
#define DEQUAN_WITH_STATS
#define DEQUAN_IMPLEMENTATION
#include "dequan.h"
#include <cstdio>
#include <vector>

using namespace dequan;

static void queens(CSP& csp, int n) {                       // test/main-test.cpp:36-49
    for (int i = 0; i < n; i++) csp.AddIntVar(0, n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, 0));
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, j - i));
            csp.AddConstraint(OpConstraint(i, j, OpConstraint::Op::NotEqual, i - j));
        }
    csp.FinalizeModel();
}

static void report(const char* name, const CSP& csp) {
    Assignment a;
    a.Reset(csp);
    b200::SolveReport rep;
    const unsigned long long count = b200::CountAll(csp, a, b200::SolveOptions(), &rep);
    Assignment e;
    e.Reset(csp);
    std::vector<std::vector<int> > all;
    b200::SolveReport erep;
    b200::EnumerateAll(csp, e, all, 100000, b200::SolveOptions(), &erep);
    bool small_refused = false;
    if (count > 1) {
        try { std::vector<std::vector<int> > few; b200::EnumerateAll(csp, e, few, count - 1); }
        catch (const b200::Error&) { small_refused = true; }
    }
    std::printf("{\"name\":\"%s\",\"count\":%llu,\"nodes\":%llu,\"enum_nodes\":%llu,\"first_in_a\":%s,\"small_refused\":%s,\"all\":[",
                name, count, (unsigned long long)rep.tree.n_nodes, (unsigned long long)erep.tree.n_nodes,
                a.IsComplete() ? "true" : "false", small_refused ? "true" : "false");
    for (size_t i = 0; i < all.size(); i++) {
        std::printf("%s[", i ? "," : "");
        for (size_t j = 0; j < all[i].size(); j++) std::printf("%s%d", j ? "," : "", all[i][j]);
        std::printf("]");
    }
    std::printf("]}\n");
}

int main() {
    try {
        { CSP csp; queens(csp, 6); report("nqueens6", csp); }
        { CSP csp; queens(csp, 8); report("nqueens8", csp); }
        {
            CSP csp;                                         // tests/enum_models.py "ordered_values"
            VarId a = csp.AddIntVar(Domain(DomainType::Values, {5, 1, 3}));
            VarId b = csp.AddIntVar(Domain(DomainType::Values, {2, 9, 4, 0}));
            VarId c = csp.AddIntVar(Domain(DomainType::Ranges, {0, 3, 7, 9}));
            csp.AddConstraint(OpConstraint(a, b, OpConstraint::Op::NotEqual, 1));
            csp.AddConstraint(OpConstraint(c, a, OpConstraint::Op::Inf, 0));
            csp.FinalizeModel();
            report("ordered_values", csp);
        }
        { CSP csp; queens(csp, 3); report("nqueens3", csp); }
    } catch (const b200::Error& e) {
        std::fprintf(stderr, "dequan::b200::Error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
