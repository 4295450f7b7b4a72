// Host-side text I/O for the drop-in CLI: at GPU rates the float-text parse of the
// input cloud (np.loadtxt, reference tutorial/tutorial_utils.py:13) and above all np.savetxt of the
// outputs (tutorial/compute_normals.py:87,91; test_c_est.py:201-210) dominate the wall time, so they
// are done here with all host threads.  Formats are byte-compatible with numpy's defaults:
// savetxt -> "%.18e", single-space delimiter, '\n' newline; loadtxt -> whitespace-separated decimal
// numbers, '#' comments, parsed as double then narrowed to float32 exactly like
// np.loadtxt(..).astype('float32').
#include <errno.h>
#include <math.h>
#include <fcntl.h>
#include <unistd.h>
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

// ---- "%.18e" of a float32-origin value without printf -------------------------------------------------------------
// x = m * 2^e2 with m < 2^24, so x * 10^p = m * 5^p * 2^(p + e2) is an exact 128-bit integer operation for every p this
// needs; the 19 digits are that integer rounded half-to-even on the exact remainder -- what printf / Python's float
// formatting produce (both round the exact binary value correctly).  Covers 2^-37 <= |x| < 2^24 (everything this
// pipeline writes except zeros, which are handled, and outliers, which fall back to snprintf).
typedef unsigned __int128 u128;
struct Pow5 {
    u128 v[40];
    Pow5() { v[0] = 1; for (int i = 1; i < 40; ++i) v[i] = v[i - 1] * 5; }
};
const Pow5 g_pow5;
const char g_digits2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263"
    "646566676869707172737475767778798081828384858687888990919293949596979899";

// returns the number of characters written (no terminator), 0 when the value is outside the fast range
inline int format_e18_f32(float xf, char* out) {
    uint32_t bits;
    memcpy(&bits, &xf, 4);
    const bool neg = (bits >> 31) != 0;
    const uint32_t ex = (bits >> 23) & 0xffu;
    uint32_t m = bits & 0x7fffffu;
    char* o = out;
    if (ex == 0 && m == 0) {   // +-0
        if (neg) *o++ = '-';
        memcpy(o, "0.000000000000000000e+00", 24);
        return (int)(o - out) + 24;
    }
    if (ex == 0 || ex == 0xffu) return 0;   // subnormal, inf, nan
    m |= 0x800000u;
    const int e2 = (int)ex - 150;           // x = m * 2^e2
    if (e2 < -60 || e2 > 0) return 0;
    const int s = -e2;
    // decimal exponent estimate from the binary one (floor(log10(x)) within one), corrected by the digit count below
    const int msb = 23 - s;                                  // x in [2^msb, 2^(msb+1))
    int k = (int)((msb * 1233) >> 12);                       // floor(msb * log10(2)) for |msb| < 2^11 (arithmetic shift)
    unsigned long long Q = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int p = 18 - k;                                // want x * 10^p in [10^18, 10^19)
        if (p < 0 || p >= 40) return 0;
        const u128 V = (u128)m * g_pow5.v[p];                // < 2^24 * 5^39: fits (5^39 < 2^91)
        const int sh = p - s;                                // x * 10^p = V * 2^sh
        u128 q;
        bool up = false;
        if (sh >= 0) {
            if (sh > 20) return 0;
            q = V << sh;
        } else {
            const int r = -sh;
            if (r >= 127) return 0;
            q = V >> r;
            const u128 rem = V & (((u128)1 << r) - 1), half = (u128)1 << (r - 1);
            up = rem > half || (rem == half && (q & 1));
        }
        const u128 lo = (u128)1000000000000000000ull, hi = lo * 10;
        if (q >= hi) { ++k; continue; }
        if (q < lo) { --k; continue; }
        Q = (unsigned long long)q + (up ? 1ull : 0ull);
        if (Q == 10000000000000000000ull) { Q = 1000000000000000000ull; ++k; }
        break;
    }
    if (Q == 0) return 0;
    if (neg) *o++ = '-';
    // 19 digits: d.dddddddddddddddddd
    char d[20];
    unsigned long long t = Q;
    for (int i = 17; i >= 1; i -= 2) {
        const unsigned two = (unsigned)(t % 100);
        t /= 100;
        d[i] = g_digits2[2 * two];
        d[i + 1] = g_digits2[2 * two + 1];
    }
    d[0] = (char)('0' + (int)t);
    *o++ = d[0];
    *o++ = '.';
    memcpy(o, d + 1, 18);
    o += 18;
    *o++ = 'e';
    int ak = k;
    if (ak < 0) { *o++ = '-'; ak = -ak; } else *o++ = '+';
    *o++ = g_digits2[2 * ak];
    *o++ = g_digits2[2 * ak + 1];
    return (int)(o - out);
}
inline int format_e18(float x, char* buf, size_t cap) {
    const int n = format_e18_f32(x, buf);
    return n ? n : snprintf(buf, cap, "%.18e", (double)x);
}
inline int format_e18(double x, char* buf, size_t cap) {
    // a double that is exactly a float32 (the f64 buffers of this pipeline mostly are not: curv / rad are true doubles)
    const float f = (float)x;
    if ((double)f == x) {
        const int n = format_e18_f32(f, buf);
        if (n) return n;
    }
    return snprintf(buf, cap, "%.18e", x);
}

// strlen of "%.18e": d.{18}e+XX is 24 characters plus the sign for every finite value with a two-digit exponent
inline int length_e18(double x) {
    const double a = fabs(x);
    if (a == 0.0 || (a >= 1e-99 && a < 1e100)) return 24 + (signbit(x) ? 1 : 0);
    char buf[64];
    return snprintf(buf, sizeof(buf), "%.18e", x);   // nan, inf, three-digit exponents
}

template <typename T>
int savetxt_impl(const char* path, const T* data, int64_t rows, int32_t cols) {
    DFB_REQUIRE(path && (data || rows == 0) && rows >= 0 && cols >= 1, DFB_E_ARG, "dfb_savetxt: bad arguments");
    const int nt = (int)((rows < 4096) ? 1 : host_threads());
    // Two passes over contiguous blocks of rows, one block per thread: (1) the byte length of the block (a number in the
    // fast range is 24 characters plus its sign; the rare others are formatted to find out), (2) after a prefix sum,
    // format into a small reusable buffer and pwrite it at the block's offset -- formatting and the copies into the
    // page cache run on all threads, and no file-sized intermediate buffer is ever allocated.
    std::vector<size_t> off(nt + 1, 0);
    std::vector<int> bad(nt, 0);
    std::vector<std::thread> th;
    auto run = [&](auto&& fn) {
        th.clear();
        for (int t = 1; t < nt; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    run([&](int t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        size_t bytes = 0;
        for (int64_t i = r0 * cols; i < r1 * cols; ++i) bytes += (size_t)length_e18((double)data[i]) + 1;
        off[t + 1] = bytes;
    });
    for (int t = 0; t < nt; ++t) off[t + 1] += off[t];
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    DFB_REQUIRE(fd >= 0, DFB_E_ARG, "dfb_savetxt: cannot open %s: %s", path, strerror(errno));
    run([&](int t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        constexpr size_t kBuf = 1 << 20;
        std::vector<char> s(kBuf + 64);
        char* o = s.data();
        size_t at = off[t];
        auto flush = [&]() {
            const char* p = s.data();
            size_t left = (size_t)(o - s.data());
            while (left > 0) {
                const ssize_t w = pwrite(fd, p, left, (off_t)at);
                if (w <= 0) { if (w < 0 && errno == EINTR) continue; bad[t] = 1; return; }
                p += w; at += (size_t)w; left -= (size_t)w;
            }
            o = s.data();
        };
        for (int64_t r = r0; r < r1 && !bad[t]; ++r) {
            const T* row = data + r * cols;
            for (int c = 0; c < cols; ++c) {
                o += format_e18(row[c], o, 40);
                *o++ = c + 1 < cols ? ' ' : '\n';
                if ((size_t)(o - s.data()) >= kBuf) flush();
            }
        }
        flush();
        if (at != off[t + 1]) bad[t] = 1;
    });
    bool ok = close(fd) == 0;
    for (int t = 0; t < nt; ++t) ok = ok && !bad[t];
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
