// Host-side text I/O for the drop-in CLI: at GPU rates the float-text parse of the
// input cloud (np.loadtxt, reference tutorial/tutorial_utils.py:13) and above all np.savetxt of the
// outputs (tutorial/compute_normals.py:87,91; test_c_est.py:201-210) dominate the wall time, so they
// are done here with all host threads.  Formats are byte-compatible with numpy's defaults:
// savetxt -> "%.18e", single-space delimiter, '\n' newline; loadtxt -> whitespace-separated decimal
// numbers, '#' comments, parsed as double then narrowed to float32 exactly like
// np.loadtxt(..).astype('float32').
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace dfb {
namespace {

int host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    if (n > 32) n = 32;
    return (int)n;
}

template <typename T>
int savetxt_impl(const char* path, const T* data, int64_t rows, int32_t cols) {
    DFB_REQUIRE(path && (data || rows == 0) && rows >= 0 && cols >= 1, DFB_E_ARG, "dfb_savetxt: bad arguments");
    const int nt = (int)((rows < 4096) ? 1 : host_threads());
    std::vector<std::string> parts(nt);
    std::vector<std::thread> th;
    auto work = [&](int t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        std::string& s = parts[t];
        s.reserve((size_t)(r1 - r0) * cols * 26 + 16);
        char buf[64];
        for (int64_t r = r0; r < r1; ++r) {
            for (int c = 0; c < cols; ++c) {
                const int n = snprintf(buf, sizeof(buf), "%.18e", (double)data[r * cols + c]);
                s.append(buf, n);
                s.push_back(c + 1 < cols ? ' ' : '\n');
            }
        }
    };
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    FILE* f = fopen(path, "wb");
    DFB_REQUIRE(f != nullptr, DFB_E_ARG, "dfb_savetxt: cannot open %s: %s", path, strerror(errno));
    bool ok = true;
    for (auto& s : parts) ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), f) == s.size());
    ok = (fclose(f) == 0) && ok;
    DFB_REQUIRE(ok, DFB_E_ARG, "dfb_savetxt: short write to %s", path);
    return DFB_OK;
}

}  // namespace
}  // namespace dfb

extern "C" int dfb_savetxt_f32(const char* path, const float* h_data, int64_t rows, int32_t cols) {
    return dfb::savetxt_impl<float>(path, h_data, rows, cols);
}

extern "C" int dfb_savetxt_f64(const char* path, const double* h_data, int64_t rows, int32_t cols) {
    return dfb::savetxt_impl<double>(path, h_data, rows, cols);
}

extern "C" int dfb_loadtxt_f32(const char* path, float* h_out, int64_t capacity, int64_t* rows_out, int32_t* cols_out) {
    using namespace dfb;
    DFB_REQUIRE(path && rows_out && cols_out, DFB_E_ARG, "dfb_loadtxt_f32: bad arguments");
    FILE* f = fopen(path, "rb");
    DFB_REQUIRE(f != nullptr, DFB_E_ARG, "dfb_loadtxt_f32: cannot open %s: %s", path, strerror(errno));
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string text((size_t)(sz > 0 ? sz : 0), '\0');
    const size_t got = sz > 0 ? fread(&text[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    DFB_REQUIRE((long)got == sz, DFB_E_ARG, "dfb_loadtxt_f32: short read of %s", path);
    text.push_back('\n');
    // line starts
    std::vector<size_t> starts;
    starts.reserve(text.size() / 24 + 16);
    size_t p = 0;
    const size_t n = text.size();
    while (p < n) {
        size_t e = p;
        while (text[e] != '\n') ++e;
        // skip blank / comment-only lines
        size_t q = p;
        while (q < e && (text[q] == ' ' || text[q] == '\t' || text[q] == '\r')) ++q;
        if (q < e && text[q] != '#') starts.push_back(p);
        text[e] = '\0';
        p = e + 1;
    }
    const int64_t rows = (int64_t)starts.size();
    // columns from the first data line
    int cols = 0;
    if (rows > 0) {
        const char* s = text.c_str() + starts[0];
        char* end = nullptr;
        for (;;) {
            while (*s == ' ' || *s == '\t' || *s == '\r' || *s == ',') ++s;
            if (*s == '\0' || *s == '#') break;
            strtod(s, &end);
            if (end == s) break;
            ++cols;
            s = end;
        }
    }
    *rows_out = rows;
    *cols_out = cols;
    if (h_out == nullptr) return DFB_OK;  // size query
    DFB_REQUIRE(cols >= 1 || rows == 0, DFB_E_ARG, "dfb_loadtxt_f32: no numeric columns in %s", path);
    DFB_REQUIRE(capacity >= rows * (int64_t)cols, DFB_E_ARG, "dfb_loadtxt_f32: output buffer too small");
    const int nt = (int)((rows < 4096) ? 1 : host_threads());
    std::vector<int> bad(nt, 0);
    std::vector<std::thread> th;
    auto work = [&](int t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        for (int64_t r = r0; r < r1; ++r) {
            const char* s = text.c_str() + starts[r];
            char* end = nullptr;
            for (int c = 0; c < cols; ++c) {
                while (*s == ' ' || *s == '\t' || *s == '\r' || *s == ',') ++s;
                const double v = strtod(s, &end);
                if (end == s) { bad[t] = 1; break; }
                h_out[r * cols + c] = (float)v;  // double -> float32, like .astype('float32')
                s = end;
            }
        }
    };
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int t = 0; t < nt; ++t) DFB_REQUIRE(!bad[t], DFB_E_ARG, "dfb_loadtxt_f32: ragged or non-numeric row in %s", path);
    return DFB_OK;
}
