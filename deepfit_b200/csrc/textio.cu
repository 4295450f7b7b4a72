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
// The same for a true double (curv / rad, compute_normals.py:79, is one): m < 2^53, so m * 5^p fits 128 bits up to p = 31,
// i.e. 1e-13 <= |x| < 1e19 -- every curvature this pipeline writes; the rest falls back.
inline int format_e18_f64(double x, char* out) {
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const bool neg = (bits >> 63) != 0;
    const int ex = (int)((bits >> 52) & 0x7ffu);
    uint64_t m = bits & ((1ull << 52) - 1);
    char* o = out;
    if (ex == 0 && m == 0) {   // +-0
        if (neg) *o++ = '-';
        memcpy(o, "0.000000000000000000e+00", 24);
        return (int)(o - out) + 24;
    }
    if (ex == 0 || ex == 0x7ff) return 0;   // subnormal, inf, nan
    m |= 1ull << 52;
    const int s = 1075 - ex;                // x = m * 2^-s
    const int msb = ex - 1023;              // x in [2^msb, 2^(msb+1))
    int k = (int)((msb * 1233) >> 12);      // floor(msb * log10(2)) within one for |msb| < 2^11 (arithmetic shift)
    unsigned long long Q = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int p = 18 - k;               // want x * 10^p in [10^18, 10^19)
        if (p < 0 || p > 31) return 0;      // 2^53 * 5^31 < 2^125
        const u128 V = (u128)m * g_pow5.v[p];
        const int sh = p - s;               // x * 10^p = V * 2^sh
        u128 q;
        bool up = false;
        if (sh >= 0) {
            if (sh > 60 || (V >> (126 - sh)) != 0) return 0;
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
    if (Q == 0 || k < -99 || k > 99) return 0;
    if (neg) *o++ = '-';
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
    int n = format_e18_f64(x, buf);
    if (n) return n;
    const float f = (float)x;   // a double that is exactly a float32: the narrower mantissa reaches further down (2^-37)
    if ((double)f == x) {
        n = format_e18_f32(f, buf);
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

// ---- loadtxt ---------------------------------------------------------------------------------------------------------
// The file is mapped, not copied; line counting and parsing both run on all host threads over byte ranges cut at line
// ends.  Numbers go through an exact fast path instead of strtod (which was 90 % of the time):
//   * up to 19 significant digits are gathered into a 64-bit integer w with a decimal exponent e;
//   * w < 2^53 and |e| <= 22: w and 10^|e| are exact doubles, so ONE IEEE multiplication / division is the correctly
//     rounded value (Clinger's fast path) -- every coordinate file with <= 15 digits;
//   * otherwise (np.savetxt's own 19-digit "%.18e" lines): the same in x87 extended precision (64-bit significand: w and
//     10^|e|, |e| <= 27, are exact), and the 64-bit result is rounded to double only when it is at least two extended ulps
//     away from a midpoint of two doubles, where the second rounding cannot differ from rounding the exact value;
//   * anything else (more digits, huge exponents, nan / inf, hex floats, near-midpoint cases) goes to strtod.
// So every value equals strtod's, i.e. np.loadtxt(..).astype('float32'), bit for bit.
namespace dfb {
namespace {

inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == '\r' || c == ','; }

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
struct Pow10L {
    long double v[28];
    Pow10L() { v[0] = 1.0L; for (int i = 1; i < 28; ++i) v[i] = v[i - 1] * 10.0L; }   // exact: 10^27 = 2^27 * 5^27, 5^27 < 2^63
};
const Pow10L kPow10L;

// Parses one number at s (s < end, not a separator).  Returns the position after it, or nullptr when there is none.
inline const char* parse_number(const char* s, const char* end, double* out) {
    const char* const s0 = s;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
    uint64_t w = 0;
    int nd = 0;            // significant digits gathered into w (leading zeros do not count)
    int dropped = 0;       // significant digits that did not fit (forces the fallback)
    int e10 = 0;
    bool any = false, fast = true;
    while (s < end && (unsigned)(*s - '0') <= 9u) {
        any = true;
        if (nd < 19) { if (nd > 0 || *s != '0') { w = w * 10 + (uint64_t)(*s - '0'); ++nd; } }
        else { ++dropped; ++e10; }
        ++s;
    }
    if (s < end && *s == '.') {
        ++s;
        while (s < end && (unsigned)(*s - '0') <= 9u) {
            any = true;
            if (nd < 19) { if (nd > 0 || *s != '0') { w = w * 10 + (uint64_t)(*s - '0'); ++nd; } --e10; }
            else ++dropped;
            ++s;
        }
    }
    if (!any) fast = false;   // nan, inf, or not a number at all
    if (any && s < end && (*s == 'e' || *s == 'E')) {
        const char* t = s + 1;
        bool eneg = false;
        if (t < end && (*t == '-' || *t == '+')) { eneg = *t == '-'; ++t; }
        if (t < end && (unsigned)(*t - '0') <= 9u) {
            int ex = 0;
            while (t < end && (unsigned)(*t - '0') <= 9u) { if (ex < 100000) ex = ex * 10 + (*t - '0'); ++t; }
            e10 += eneg ? -ex : ex;
            s = t;
        }
    }
    // a token must end at a separator, the line end or a comment; anything else (hex floats, "1.5abc") is strtod's call
    if (s < end && !(is_sep(*s) || *s == '\n' || *s == '#' || *s == '\0')) fast = false;
    if (fast && dropped == 0) {
        if (w == 0) { *out = neg ? -0.0 : 0.0; return s; }
        if (w < (1ull << 53) && e10 >= -22 && e10 <= 22) {
            const double d = (double)w;
            const double r = e10 < 0 ? d / kPow10[-e10] : d * kPow10[e10];
            *out = neg ? -r : r;
            return s;
        }
#if (defined(__x86_64__) || defined(__i386__)) && defined(__LDBL_MANT_DIG__) && __LDBL_MANT_DIG__ == 64   // x87 extended only
        if (e10 >= -27 && e10 <= 27) {
            const long double d = (long double)w;                  // exact: 64-bit significand
            const long double r = e10 < 0 ? d / kPow10L.v[-e10] : d * kPow10L.v[e10];
            uint64_t sig;
            memcpy(&sig, &r, 8);                                   // x87 extended: the first 8 bytes are the significand
            const uint64_t low = sig & 0x7ffull;                   // the bits below a double's last place
            if (r > 1e-290L && r < 1e290L && (low < 0x3feull || low > 0x402ull)) {
                const double rd = (double)r;
                *out = neg ? -rd : rd;
                return s;
            }
        }
#endif
    }
    // fallback: strtod on a NUL-terminated copy of the token
    char buf[128];
    const char* t = s0;
    while (t < end && !(is_sep(*t) || *t == '\n' || *t == '\0')) ++t;
    size_t len = (size_t)(t - s0);
    std::string big;
    const char* src;
    if (len < sizeof(buf)) { memcpy(buf, s0, len); buf[len] = '\0'; src = buf; }
    else { big.assign(s0, len); src = big.c_str(); }
    char* ep = nullptr;
    const double v = strtod(src, &ep);
    if (ep == src) return nullptr;
    *out = v;
    return s0 + (ep - src);
}

// first byte of the next line at or after p (end when there is none)
inline const char* next_line(const char* p, const char* end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl + 1 : end;
}
// a data line has a first non-blank character that is not '#'
inline bool is_data_line(const char* p, const char* line_end) {
    while (p < line_end && (is_sep(*p))) ++p;
    return p < line_end && *p != '#' && *p != '\n';
}

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool mapped = false;
    std::string copy;   // fallback when mmap is not available (pipes, odd file systems)
    ~Mapped();
};

}  // namespace
}  // namespace dfb

#include <sys/mman.h>
#include <sys/stat.h>

dfb::Mapped::~Mapped() {
    if (mapped) munmap((void*)p, n);
    if (fd >= 0) close(fd);
}

extern "C" int dfb_loadtxt_f32(const char* path, float* h_out, int64_t capacity, int64_t* rows_out, int32_t* cols_out) {
    using namespace dfb;
    DFB_REQUIRE(path && rows_out && cols_out, DFB_E_ARG, "dfb_loadtxt_f32: bad arguments");
    Mapped m;
    m.fd = open(path, O_RDONLY);
    DFB_REQUIRE(m.fd >= 0, DFB_E_ARG, "dfb_loadtxt_f32: cannot open %s: %s", path, strerror(errno));
    struct stat st;
    DFB_REQUIRE(fstat(m.fd, &st) == 0, DFB_E_ARG, "dfb_loadtxt_f32: cannot stat %s: %s", path, strerror(errno));
    m.n = (size_t)st.st_size;
    if (m.n > 0) {
        void* a = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
        if (a != MAP_FAILED) {
            m.p = (const char*)a;
            m.mapped = true;
            madvise(a, m.n, MADV_SEQUENTIAL);
        } else {
            m.copy.resize(m.n);
            size_t got = 0;
            while (got < m.n) {
                const ssize_t r = pread(m.fd, &m.copy[got], m.n - got, (off_t)got);
                if (r <= 0) break;
                got += (size_t)r;
            }
            DFB_REQUIRE(got == m.n, DFB_E_ARG, "dfb_loadtxt_f32: short read of %s", path);
            m.p = m.copy.data();
        }
    }
    const char* const base = m.p;
    const char* const end = m.p + m.n;
    // byte ranges cut at line starts, one per thread
    const int nt = (int)((m.n < (1u << 20)) ? 1 : host_threads());
    std::vector<const char*> cut(nt + 1, end);
    cut[0] = base;
    for (int t = 1; t < nt; ++t) {
        const char* guess = base + m.n * (size_t)t / (size_t)nt;
        cut[t] = guess <= cut[t - 1] ? cut[t - 1] : next_line(guess, end);
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    // pass 1: data lines per range
    std::vector<int64_t> count(nt, 0);
    auto count_range = [&](int t) {
        int64_t c = 0;
        for (const char* p = cut[t]; p < cut[t + 1];) {
            const char* ne = next_line(p, end);
            if (is_data_line(p, ne)) ++c;
            p = ne;
        }
        count[t] = c;
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(count_range, t);
        count_range(0);
        for (auto& x : th) x.join();
    }
    int64_t rows = 0;
    std::vector<int64_t> first(nt, 0);
    for (int t = 0; t < nt; ++t) { first[t] = rows; rows += count[t]; }
    // columns from the first data line
    int cols = 0;
    bool first_line_bad = false;
    for (const char* p = base; p < end;) {
        const char* ne = next_line(p, end);
        if (is_data_line(p, ne)) {
            const char* s = p;
            for (;;) {
                while (s < ne && is_sep(*s)) ++s;
                if (s >= ne || *s == '\n' || *s == '#' || *s == '\0') break;
                double v;
                const char* e = parse_number(s, ne, &v);
                if (!e || e == s) { first_line_bad = true; break; }
                ++cols;
                s = e;
            }
            break;
        }
        p = ne;
    }
    DFB_REQUIRE(!first_line_bad, DFB_E_ARG, "dfb_loadtxt_f32: non-numeric token in the first data row of %s", path);
    *rows_out = rows;
    *cols_out = cols;
    if (h_out == nullptr) return DFB_OK;  // size query
    DFB_REQUIRE(cols >= 1 || rows == 0, DFB_E_ARG, "dfb_loadtxt_f32: no numeric columns in %s", path);
    DFB_REQUIRE(capacity >= rows * (int64_t)cols, DFB_E_ARG, "dfb_loadtxt_f32: output buffer too small");
    // pass 2: parse
    std::vector<int> bad(nt, 0);
    auto parse_range = [&](int t) {
        int64_t r = first[t];
        for (const char* p = cut[t]; p < cut[t + 1];) {
            const char* ne = next_line(p, end);
            if (is_data_line(p, ne)) {
                const char* s = p;
                for (int c = 0; c < cols; ++c) {
                    while (s < ne && is_sep(*s)) ++s;
                    double v;
                    const char* e = (s < ne && *s != '\n' && *s != '#') ? parse_number(s, ne, &v) : nullptr;
                    if (!e || e == s) { bad[t] = 1; break; }
                    h_out[r * cols + c] = (float)v;  // double -> float32, like .astype('float32')
                    s = e;
                }
                while (s < ne && is_sep(*s)) ++s;
                if (s < ne && *s != '\n' && *s != '#' && *s != '\0') bad[t] = 1;   // more columns than the first row
                ++r;
            }
            p = ne;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(parse_range, t);
        parse_range(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < nt; ++t) DFB_REQUIRE(!bad[t], DFB_E_ARG, "dfb_loadtxt_f32: ragged or non-numeric row in %s", path);
    return DFB_OK;
}
