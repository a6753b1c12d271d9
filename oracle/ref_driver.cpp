// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Drives the UNMODIFIED reference header (/root/reference/dequan.h, passed on the
// compiler command line with -I, never copied into this repo) so that the CPU
// restatement (oracle/dq_oracle.c) and the CUDA engine can be pinned against the
// real thing.  Built by oracle/Makefile into oracle/_ref/dequan_ref (git-ignored).
//
// dequan.h has no all-solutions mode, no node budget and no model file format
// (SURVEY.md §0), so this driver adds, using only the reference's own extension
// point `dequan::Constraint` (dequan.h:134-148):
//   * CountingConstraint  — linked LAST to the LAST variable of assign_order; its
//     Evaluate() counts a solution and answers Failed, so ForwardCheckingStep walks
//     the whole tree and stats.assigned_vars is the full-tree node count (§8c).
//   * BudgetConstraint    — linked FIRST to every variable; counts Evaluate() calls
//     (= AssignVar calls = nodes) and throws once the count exceeds the budget.
//   * TableConstraint     — a user-defined binary check-only constraint (tabulated
//     allowed pairs; default AplyArcConsistency, dequan.h:147).
// and a tiny text model format (".dqm", see read_model) shared with the tests.
//
// Commands (all print one JSON object per solved model on stdout):
//   dequan_ref tests                      reference scenarios of test/main-test.cpp
//   dequan_ref solve  first|count BUDGET  < model.dqm      (stream of >=1 models)
//   dequan_ref nqueens N first|count [THREADS]             (depth-1 split if THREADS>1)
//   dequan_ref sudoku FILE boxes|noboxes THREADS [LIMIT]   81-char lines, '0'/'.' blank
//   dequan_ref color  FILE K BUDGET THREADS                binary graphs, see read_graphs

#include <climits>
#include <new>
#include <utility>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <thread>
#include <atomic>
#include <iostream>
#include <sstream>
#include <fstream>
#include <stdexcept>

#define DEQUAN_USE_STDVECTOR
#define DEQUAN_WITH_STATS
#define DEQUAN_IMPLEMENTATION
#include "dequan.h"   // resolved through -I/root/reference by oracle/Makefile

using namespace dequan;
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------
struct SolveCounters {
    u64 solutions = 0;
    u64 budget = 0;          // 0 = unlimited
    u64 evals = 0;           // BudgetConstraint evaluate calls
    std::vector<int> first;  // inst_vars snapshot at first counted solution
    u64 keep_all = 0;        // `enumerate`: snapshot up to this many counted solutions, in visiting order
    std::vector<std::vector<int> > all;
};

struct BudgetExceeded {};

struct CountingConstraint : public Constraint {
    CountingConstraint(VarId last, SolveCounters* c) : last_vid(last), ctr(c) {
        static_assert(sizeof(CountingConstraint) <= GenericConstraint::MAX_CONSTRAINT_SIZE, "");
    }
    virtual void LinkVars(Array<Var>& vars) { vars[last_vid].linked_constraints.push_back(this); }
    virtual Eval Evaluate(const Array<InstVar>& inst_vars, VarId) {
        if (ctr->solutions == 0) {
            ctr->first.resize(inst_vars.size());
            for (size_t i = 0; i < inst_vars.size(); i++) ctr->first[i] = inst_vars[i].value;
        }
        if (ctr->all.size() < ctr->keep_all) {
            std::vector<int> snap(inst_vars.size());
            for (size_t i = 0; i < inst_vars.size(); i++) snap[i] = inst_vars[i].value;
            ctr->all.push_back(snap);
        }
        ctr->solutions++;
        return Eval::Failed;
    }
    VarId last_vid;
    SolveCounters* ctr;
};

struct BudgetConstraint : public Constraint {
    explicit BudgetConstraint(SolveCounters* c) : ctr(c) {
        static_assert(sizeof(BudgetConstraint) <= GenericConstraint::MAX_CONSTRAINT_SIZE, "");
    }
    virtual void LinkVars(Array<Var>& vars) {
        for (size_t i = 0; i < vars.size(); i++) vars[i].linked_constraints.push_back(this);
    }
    virtual Eval Evaluate(const Array<InstVar>&, VarId) {
        if (++ctr->evals > ctr->budget) throw BudgetExceeded();
        return Eval::Passed;
    }
    SolveCounters* ctr;
};

struct TableConstraint : public Constraint {
    TableConstraint(VarId a, VarId b, const std::vector<int>* p) : v0(a), v1(b), pairs(p) {
        static_assert(sizeof(TableConstraint) <= GenericConstraint::MAX_CONSTRAINT_SIZE, "");
    }
    virtual void LinkVars(Array<Var>& vars) {
        vars[v0].linked_constraints.push_back(this);
        vars[v1].linked_constraints.push_back(this);
    }
    virtual Eval Evaluate(const Array<InstVar>& iv, VarId) {
        int a = iv[v0].value, b = iv[v1].value;
        if (a == InstVar::UNASSIGNED || b == InstVar::UNASSIGNED) return Eval::NA;
        for (size_t i = 0; i + 1 < pairs->size(); i += 2)
            if ((*pairs)[i] == a && (*pairs)[i + 1] == b) return Eval::Passed;
        return Eval::Failed;
    }
    VarId v0, v1;
    const std::vector<int>* pairs;
};

// ---------------------------------------------------------------------------------
// Model text format:
//   dqm <n_vars> <n_cons>
//   V <n> v1..vn | R <n> m0 M0 ...                       (n_vars lines)
//   op v0 v1 OPCODE off | eq v0 v1 | alldiff n v.. | orrange v0 v1 min max
//   | table v0 v1 npairs a0 b0 ...                       (n_cons lines)
struct ModelText {
    struct Dom { int type; std::vector<int> vals; };
    struct Con { std::string kind; std::vector<int> data; };
    std::vector<Dom> doms;
    std::vector<Con> cons;
};

static bool read_model(std::istream& in, ModelText& m) {
    std::string tag;
    if (!(in >> tag)) return false;
    if (tag != "dqm") throw std::runtime_error("bad model header: " + tag);
    int nv, nc;
    in >> nv >> nc;
    m.doms.assign(nv, ModelText::Dom());
    m.cons.assign(nc, ModelText::Con());
    for (int i = 0; i < nv; i++) {
        std::string t; int n;
        in >> t >> n;
        m.doms[i].type = (t == "R") ? 1 : 0;
        m.doms[i].vals.resize(n);
        for (int j = 0; j < n; j++) in >> m.doms[i].vals[j];
    }
    for (int i = 0; i < nc; i++) {
        std::string k; in >> k;
        m.cons[i].kind = k;
        int n = 0;
        if (k == "op") n = 4; else if (k == "eq") n = 2; else if (k == "orrange") n = 4;
        else if (k == "alldiff") { in >> n; }
        else if (k == "table") { int a, b, np; in >> a >> b >> np; m.cons[i].data.push_back(a); m.cons[i].data.push_back(b); n = 2 * np; }
        else throw std::runtime_error("bad constraint kind: " + k);
        for (int j = 0; j < n; j++) { int x; in >> x; m.cons[i].data.push_back(x); }
    }
    if (!in) throw std::runtime_error("truncated model");
    return true;
}

struct BuiltModel {
    CSP csp;
    std::vector<std::vector<int> > table_pairs;  // owned storage for TableConstraint
};

// budget constraint first, model constraints in file order, counting constraint last.
static void build_model(const ModelText& m, BuiltModel& b, SolveCounters& ctr, bool count_all) {
    for (size_t i = 0; i < m.doms.size(); i++)
        b.csp.AddIntVar(Domain(m.doms[i].type ? DomainType::Ranges : DomainType::Values, m.doms[i].vals));
    if (ctr.budget) b.csp.AddConstraint(BudgetConstraint(&ctr));
    b.table_pairs.reserve(m.cons.size());
    for (size_t i = 0; i < m.cons.size(); i++) {
        const std::vector<int>& d = m.cons[i].data;
        const std::string& k = m.cons[i].kind;
        if (k == "op") b.csp.AddConstraint(OpConstraint(d[0], d[1], (OpConstraint::Op)d[2], d[3]));
        else if (k == "eq") b.csp.AddConstraint(EqualityConstraint(d[0], d[1]));
        else if (k == "orrange") b.csp.AddConstraint(OrRangeConstraint(d[0], d[1], d[2], d[3]));
        else if (k == "alldiff") b.csp.AddConstraint(AllDifferentConstraint(d));
        else if (k == "table") {
            b.table_pairs.push_back(std::vector<int>(d.begin() + 2, d.end()));
            b.csp.AddConstraint(TableConstraint(d[0], d[1], &b.table_pairs.back()));
        }
    }
    if (count_all && !m.doms.empty()) {
        Assignment probe;
        probe.Reset(b.csp);  // Reset only reads vars.size()/domains (dequan.h:365-395)
        b.csp.AddConstraint(CountingConstraint(probe.assign_order.back(), &ctr));
    }
    b.csp.FinalizeModel();
}

struct SolveOut {
    const char* status;
    u64 solutions, nodes, applied_arcs, validated;
    std::vector<int> first;
    std::vector<int> order;
    double seconds;
};

static SolveOut run_solve(const ModelText& m, bool count_all, u64 budget) {
    SolveCounters ctr;
    ctr.budget = budget;
    BuiltModel b;
    build_model(m, b, ctr, count_all);
    Assignment a;
    a.Reset(b.csp);
    SolveOut o;
    o.order.assign(a.assign_order.begin(), a.assign_order.end());
    bool ok = false, busted = false;
    auto t0 = std::chrono::steady_clock::now();
    try { ok = b.csp.ForwardCheckingStep(a); } catch (const BudgetExceeded&) { busted = true; }
    auto t1 = std::chrono::steady_clock::now();
    o.seconds = std::chrono::duration<double>(t1 - t0).count();
    o.nodes = a.stats.assigned_vars;
    o.applied_arcs = a.stats.applied_arcs;
    o.validated = a.stats.validated_constraints;
    o.solutions = count_all ? ctr.solutions : (ok ? 1 : 0);
    if (count_all) o.first = ctr.first;
    else if (ok) { o.first.resize(m.doms.size()); for (size_t i = 0; i < m.doms.size(); i++) o.first[i] = a.GetInstVarValue((int)i); }
    o.status = busted ? "budget" : (o.solutions ? "sat" : "unsat");
    return o;
}

static std::string to_json(const SolveOut& o) {
    std::ostringstream s;
    s << "{\"status\":\"" << o.status << "\",\"solutions\":" << o.solutions << ",\"nodes\":" << o.nodes
      << ",\"applied_arcs\":" << o.applied_arcs << ",\"validated_constraints\":" << o.validated << ",\"first\":";
    if (o.first.empty()) s << "null";
    else { s << "["; for (size_t i = 0; i < o.first.size(); i++) s << (i ? "," : "") << o.first[i]; s << "]"; }
    s << ",\"order\":[";
    for (size_t i = 0; i < o.order.size(); i++) s << (i ? "," : "") << o.order[i];
    s << "],\"seconds\":" << o.seconds << "}";
    return s.str();
}
static void print_out(const SolveOut& o) { std::cout << to_json(o) << std::endl; }

// ---------------------------------------------------------------------------------
// Model builders that follow test/main-test.cpp
static ModelText nqueens_model(int n, const std::vector<int>& fixed /* singleton domains for vars 0..k-1 */) {
    ModelText m;
    m.doms.resize(n);
    for (int i = 0; i < n; i++) { m.doms[i].type = 1; m.doms[i].vals = {0, n}; }   // AddIntVar(0,N) main-test.cpp:36
    // prefix split (§8c): fixed vars keep ids 0..k-1 and, being the smallest domains, stay first in assign_order
    for (size_t i = 0; i < fixed.size(); i++) { m.doms[i].type = 0; m.doms[i].vals = {fixed[i]}; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {                                           // main-test.cpp:39-48
            m.cons.push_back({"op", {i, j, 1, 0}});
            m.cons.push_back({"op", {i, j, 1, j - i}});
            m.cons.push_back({"op", {i, j, 1, i - j}});
        }
    return m;
}

static ModelText sudoku_model(const char* cells /*81 chars*/, bool boxes, bool alldiff) {
    ModelText m;
    m.doms.resize(81);
    for (int i = 0; i < 81; i++) {
        int g = (cells[i] >= '1' && cells[i] <= '9') ? cells[i] - '0' : 0;
        if (g) { m.doms[i].type = 0; m.doms[i].vals = {g}; }       // AddFixedVar main-test.cpp:125
        else   { m.doms[i].type = 1; m.doms[i].vals = {1, 10}; }   // AddIntVar(1,10) main-test.cpp:121
    }
    std::vector<std::vector<int> > groups;
    for (int r = 0; r < 9; r++) { std::vector<int> g; for (int c = 0; c < 9; c++) g.push_back(r * 9 + c); groups.push_back(g); }
    for (int c = 0; c < 9; c++) { std::vector<int> g; for (int r = 0; r < 9; r++) g.push_back(r * 9 + c); groups.push_back(g); }
    if (boxes)
        for (int b = 0; b < 9; b++) { std::vector<int> g; for (int k = 0; k < 9; k++) g.push_back((b / 3 * 3 + k / 3) * 9 + (b % 3 * 3 + k % 3)); groups.push_back(g); }
    if (alldiff) { for (auto& g : groups) m.cons.push_back({"alldiff", g}); }
    else {
        // all-different expanded to binary !=, one per unordered peer pair (810 with boxes)
        std::vector<char> seen(81 * 81, 0);
        for (auto& g : groups)
            for (size_t a = 0; a < g.size(); a++)
                for (size_t b = a + 1; b < g.size(); b++) {
                    int u = g[a], v = g[b]; if (u > v) std::swap(u, v);
                    if (seen[u * 81 + v]) continue; seen[u * 81 + v] = 1;
                    m.cons.push_back({"op", {u, v, 1, 0}});
                }
    }
    return m;
}

// ---------------------------------------------------------------------------------
static int cmd_tests() {
    // OpInequalityTest, main-test.cpp:187-233
    {
        ModelText m; m.doms.resize(4);
        m.doms[0] = {1, {0, 10}}; m.doms[1] = {1, {0, 10}}; m.doms[2] = {0, {6}}; m.doms[3] = {0, {5}};
        m.cons.push_back({"op", {0, 2, 5, 0}}); m.cons.push_back({"op", {0, 3, 2, 0}});
        m.cons.push_back({"op", {1, 2, 4, 0}}); m.cons.push_back({"op", {1, 3, 3, 0}});
        std::cout << "{\"test\":\"OpInequalityTest\",\"result\":" << to_json(run_solve(m, false, 0)) << "}" << std::endl;
    }
    { std::cout << "{\"test\":\"NQueensTest8\",\"result\":" << to_json(run_solve(nqueens_model(8, std::vector<int>()), false, 0)) << "}" << std::endl; }
    const char* grid = "003020600900305001001806400008102900700000008006708200002609500800203009005010300";  // main-test.cpp:92-105
    { std::cout << "{\"test\":\"SudokuTest_rows_cols_alldiff\",\"result\":" << to_json(run_solve(sudoku_model(grid, false, true), false, 0)) << "}" << std::endl; }
    { std::cout << "{\"test\":\"Sudoku_rows_cols_binary\",\"result\":" << to_json(run_solve(sudoku_model(grid, false, false), false, 0)) << "}" << std::endl; }
    { std::cout << "{\"test\":\"Sudoku_boxes_alldiff\",\"result\":" << to_json(run_solve(sudoku_model(grid, true, true), false, 0)) << "}" << std::endl; }
    { std::cout << "{\"test\":\"Sudoku_boxes_binary\",\"result\":" << to_json(run_solve(sudoku_model(grid, true, false), false, 0)) << "}" << std::endl; }
    return 0;
}

// enumerate CAP : models on stdin as for `solve`; one JSON line per model with every counted solution in visiting order
static int cmd_enumerate(int argc, char** argv) {
    if (argc < 3) return 2;
    u64 cap = strtoull(argv[2], 0, 10);
    ModelText m;
    while (read_model(std::cin, m)) {
        SolveCounters ctr;
        ctr.keep_all = cap;
        BuiltModel b;
        build_model(m, b, ctr, true);
        Assignment a;
        a.Reset(b.csp);
        b.csp.ForwardCheckingStep(a);
        std::cout << "{\"solutions\":" << ctr.solutions << ",\"nodes\":" << a.stats.assigned_vars << ",\"all\":[";
        for (size_t i = 0; i < ctr.all.size(); i++) {
            std::cout << (i ? "," : "") << "[";
            for (size_t j = 0; j < ctr.all[i].size(); j++) std::cout << (j ? "," : "") << ctr.all[i][j];
            std::cout << "]";
        }
        std::cout << "]}" << std::endl;
    }
    return 0;
}

static int cmd_solve(int argc, char** argv) {
    if (argc < 4) return 2;
    bool count_all = !strcmp(argv[2], "count");
    u64 budget = strtoull(argv[3], 0, 10);
    ModelText m;
    while (read_model(std::cin, m)) print_out(run_solve(m, count_all, budget));
    return 0;
}

template <class F>
static void parallel_for(long n, int threads, F f) {
    std::atomic<long> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&]() { for (;;) { long i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& t : th) t.join();
}

// nqueens N first|count [THREADS [p0 p1 ...]] : with THREADS>1 (count only) the tree below the fixed
// prefix p0.. is split on the next variable's values, one task per value, and the results summed.
static int cmd_nqueens(int argc, char** argv) {
    if (argc < 4) return 2;
    int n = atoi(argv[2]);
    bool count_all = !strcmp(argv[3], "count");
    int threads = argc > 4 ? atoi(argv[4]) : 1;
    std::vector<int> prefix;
    for (int i = 5; i < argc; i++) prefix.push_back(atoi(argv[i]));
    auto t0 = std::chrono::steady_clock::now();
    if (threads <= 1 || !count_all || (int)prefix.size() >= n) {
        SolveOut o = run_solve(nqueens_model(n, prefix), count_all, 0);
        print_out(o);
        return 0;
    }
    std::vector<SolveOut> outs(n);
    parallel_for(n, threads, [&](long v) { std::vector<int> p = prefix; p.push_back((int)v); outs[v] = run_solve(nqueens_model(n, p), true, 0); });
    auto t1 = std::chrono::steady_clock::now();
    SolveOut sum = outs[0];
    for (int v = 1; v < n; v++) {
        sum.solutions += outs[v].solutions; sum.nodes += outs[v].nodes;
        sum.applied_arcs += outs[v].applied_arcs; sum.validated += outs[v].validated;
        if (sum.first.empty()) sum.first = outs[v].first;
    }
    sum.status = sum.solutions ? "sat" : "unsat";
    sum.seconds = std::chrono::duration<double>(t1 - t0).count();
    print_out(sum);
    return 0;
}

static int cmd_sudoku(int argc, char** argv) {
    if (argc < 5) return 2;
    std::ifstream f(argv[2]);
    bool boxes = !strcmp(argv[3], "boxes");
    int threads = atoi(argv[4]);
    long limit = argc > 5 ? atol(argv[5]) : -1;
    bool quiet = argc > 6 && !strcmp(argv[6], "quiet");
    std::vector<std::string> lines; std::string s;
    while (std::getline(f, s)) { if (s.size() >= 81) lines.push_back(s.substr(0, 81)); if (limit >= 0 && (long)lines.size() >= limit) break; }
    std::vector<SolveOut> outs(lines.size());
    auto t0 = std::chrono::steady_clock::now();
    parallel_for((long)lines.size(), threads, [&](long i) { outs[i] = run_solve(sudoku_model(lines[i].c_str(), boxes, false), false, 0); });
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    u64 nodes = 0; double solve_s = 0;
    for (auto& o : outs) { nodes += o.nodes; solve_s += o.seconds; if (!quiet) print_out(o); }
    std::cout << "{\"summary\":\"sudoku\",\"puzzles\":" << lines.size() << ",\"threads\":" << threads << ",\"nodes\":" << nodes
              << ",\"wall_seconds\":" << wall << ",\"solve_seconds_sum\":" << solve_s << "}" << std::endl;
    return 0;
}

// graphs file: text, one instance per line: "n m u0 v0 u1 v1 ..."
static int cmd_color(int argc, char** argv) {
    if (argc < 6) return 2;
    std::ifstream f(argv[2]);
    int k = atoi(argv[3]);
    u64 budget = strtoull(argv[4], 0, 10);
    int threads = atoi(argv[5]);
    bool quiet = argc > 6 && !strcmp(argv[6], "quiet");
    std::vector<ModelText> models; std::string s;
    while (std::getline(f, s)) {
        std::istringstream is(s); int n, m; if (!(is >> n >> m)) continue;
        ModelText mt; mt.doms.resize(n);
        for (int i = 0; i < n; i++) mt.doms[i] = {1, {0, k}};
        for (int e = 0; e < m; e++) { int u, v; is >> u >> v; mt.cons.push_back({"op", {u, v, 1, 0}}); }
        models.push_back(mt);
    }
    std::vector<SolveOut> outs(models.size());
    auto t0 = std::chrono::steady_clock::now();
    parallel_for((long)models.size(), threads, [&](long i) { outs[i] = run_solve(models[i], false, budget); });
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    u64 nodes = 0; double solve_s = 0;
    for (auto& o : outs) { nodes += o.nodes; solve_s += o.seconds; if (!quiet) print_out(o); }
    std::cout << "{\"summary\":\"color\",\"instances\":" << models.size() << ",\"threads\":" << threads << ",\"nodes\":" << nodes
              << ",\"wall_seconds\":" << wall << ",\"solve_seconds_sum\":" << solve_s << "}" << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    std::ios::sync_with_stdio(false);
    try {
        if (argc < 2) { fprintf(stderr, "usage: dequan_ref tests|solve|enumerate|nqueens|sudoku|color ...\n"); return 2; }
        if (!strcmp(argv[1], "tests")) return cmd_tests();
        if (!strcmp(argv[1], "solve")) return cmd_solve(argc, argv);
        if (!strcmp(argv[1], "enumerate")) return cmd_enumerate(argc, argv);
        if (!strcmp(argv[1], "nqueens")) return cmd_nqueens(argc, argv);
        if (!strcmp(argv[1], "sudoku")) return cmd_sudoku(argc, argv);
        if (!strcmp(argv[1], "color")) return cmd_color(argc, argv);
    } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 2;
}
