// dq_formats.cpp — host-side readers for the on-disk instance formats of the batch entry points
// (SURVEY.md §8f-3): 81-character Sudoku lines and DIMACS .col graphs.  Pure parsing: the solve
// itself is dq_solve_batch_cells / dq_solve_batch_graphs.
#include <cstdio>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>

#include "dequan_b200.h"

namespace dq { void set_last_error(const std::string& s); }

extern "C" {

int dq_parse_sudoku_lines(const char* text, size_t len, uint8_t* cells, int64_t cap, int64_t* n_out) {
    if (!text || !n_out || (cap > 0 && !cells)) { dq::set_last_error("null argument"); return DQ_ERR_INVALID; }
    int64_t n = 0;
    size_t i = 0;
    while (i < len) {
        size_t e = i;
        while (e < len && text[e] != '\n') e++;
        size_t a = i, b = e;
        while (a < b && isspace((unsigned char)text[a])) a++;
        while (b > a && isspace((unsigned char)text[b - 1])) b--;
        if (b > a && text[a] != '#') {
            if (b - a != 81) { dq::set_last_error("line " + std::to_string(n + 1) + ": expected 81 characters, got " + std::to_string(b - a)); return DQ_ERR_INVALID; }
            if (n >= cap) { dq::set_last_error("more puzzles than the output buffer holds"); return DQ_ERR_NOMEM; }
            for (size_t k = 0; k < 81; k++) {
                const char ch = text[a + k];
                uint8_t v;
                if (ch >= '1' && ch <= '9') v = (uint8_t)(ch - '0');
                else if (ch == '0' || ch == '.' || ch == '_' || ch == '*') v = 0;
                else { dq::set_last_error(std::string("line ") + std::to_string(n + 1) + ": bad character '" + ch + "'"); return DQ_ERR_INVALID; }
                cells[n * 81 + (int64_t)k] = v;
            }
            n++;
        }
        i = e + 1;
    }
    *n_out = n;
    return DQ_OK;
}

int dq_parse_dimacs_col(const char* text, size_t len, int32_t* n_vertices, uint8_t* edges, int64_t cap_edges, int64_t* n_edges) {
    if (!text || !n_vertices || !n_edges || (cap_edges > 0 && !edges)) { dq::set_last_error("null argument"); return DQ_ERR_INVALID; }
    int nv = -1;
    int64_t m = 0;
    size_t i = 0;
    int line_no = 0;
    while (i < len) {
        size_t e = i;
        while (e < len && text[e] != '\n') e++;
        line_no++;
        std::string line(text + i, e - i);
        i = e + 1;
        size_t a = 0;
        while (a < line.size() && isspace((unsigned char)line[a])) a++;
        if (a == line.size() || line[a] == 'c') continue;
        if (line[a] == 'p') {
            char kind[32];
            long n = 0, me = 0;
            if (sscanf(line.c_str() + a, "p %31s %ld %ld", kind, &n, &me) != 3 || n < 1) { dq::set_last_error("bad problem line"); return DQ_ERR_INVALID; }
            if (n > 254) { dq::set_last_error("more than 254 vertices"); return DQ_ERR_UNSUPPORTED; }
            nv = (int)n;
        } else if (line[a] == 'e') {
            long u = 0, v = 0;
            if (nv < 0 || sscanf(line.c_str() + a, "e %ld %ld", &u, &v) != 2 || u < 1 || v < 1 || u > nv || v > nv || u == v) {
                dq::set_last_error("line " + std::to_string(line_no) + ": bad edge"); return DQ_ERR_INVALID;
            }
            if (m >= cap_edges) { dq::set_last_error("more edges than the output buffer holds"); return DQ_ERR_NOMEM; }
            edges[2 * m] = (uint8_t)(u - 1);
            edges[2 * m + 1] = (uint8_t)(v - 1);
            m++;
        } else { dq::set_last_error("line " + std::to_string(line_no) + ": unknown record"); return DQ_ERR_INVALID; }
    }
    if (nv < 0) { dq::set_last_error("no problem line"); return DQ_ERR_INVALID; }
    *n_vertices = nv;
    *n_edges = m;
    return DQ_OK;
}

}  // extern "C"
