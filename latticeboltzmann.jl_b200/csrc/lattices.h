// Compile-time velocity sets of the seven quadratures (population order == column order of
// `abscissae` in /root/reference/src/quadratures/D2Q*.jl; indices 0-based) and the host-side
// weight tables.  Weights are evaluated with the same double operations the reference uses so
// the values are bit-identical to Julia's (checked against the oracle in tests/test_abi.py).
#pragma once
#include <cmath>
#include "../../include/lbm_b200.h"

#if defined(__CUDACC__)
#define LBM_HD __host__ __device__
#else
#define LBM_HD
#endif

namespace lbm {

template <int ID>
struct Lat;

// opposite(q, idx), 1-based rules restated 0-based.
// generic: src/quadratures.jl:11-19
LBM_HD constexpr int opp_generic(int i) { return i == 0 ? 0 : ((i + 1) % 2 == 0 ? i + 1 : i - 1); }

template <>
struct Lat<LBM_D2Q4> {  // src/quadratures/D2Q4.jl:16-31
    static constexpr int Q = 4, EQ_ORDER = 1, N = 1, H = 1;
    LBM_HD static constexpr int cx(int i) { constexpr int t[Q] = {1, 0, -1, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[Q] = {0, 1, 0, -1}; return t[i]; }
    LBM_HD static constexpr int opp(int i) { return i <= 1 ? i + 2 : i - 2; }
    static constexpr bool UNIT_PRESSURE = true;  // velocity_distribution_function/quadratures.jl:127
};
template <>
struct Lat<LBM_D2Q5> {  // src/quadratures/D2Q5.jl:17-48
    static constexpr int Q = 5, EQ_ORDER = 1, N = 1, H = 1;
    LBM_HD static constexpr int cx(int i) { constexpr int t[Q] = {0, 1, 0, -1, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[Q] = {0, 0, 1, 0, -1}; return t[i]; }
    LBM_HD static constexpr int opp(int i) { return i == 0 ? 0 : (i <= 2 ? i + 2 : i - 2); }
    static constexpr bool UNIT_PRESSURE = true;  // quadratures/D2Q5.jl:48
};
template <>
struct Lat<LBM_D2Q9> {  // src/quadratures/D2Q9.jl:20-38
    static constexpr int Q = 9, EQ_ORDER = 2, N = 2, H = 1;
    LBM_HD static constexpr int cx(int i) { constexpr int t[Q] = {0, -1, -1, -1, 0, 1, 1, 1, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[Q] = {0, 1, 0, -1, -1, -1, 0, 1, 1}; return t[i]; }
    LBM_HD static constexpr int opp(int i) { return i == 0 ? 0 : (i <= 4 ? i + 4 : i - 4); }
    static constexpr bool UNIT_PRESSURE = false;
};
template <>
struct Lat<LBM_D2Q13> {  // src/quadratures/D2Q13.jl:10-29
    static constexpr int Q = 13, EQ_ORDER = 2, N = 2, H = 2;
    LBM_HD static constexpr int cx(int i) { constexpr int t[Q] = {0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[Q] = {0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, -2, 2}; return t[i]; }
    LBM_HD static constexpr int opp(int i) { return opp_generic(i); }
    static constexpr bool UNIT_PRESSURE = false;
};
template <>
struct Lat<LBM_D2Q17> {  // src/quadratures/D2Q17.jl:21-57
    static constexpr int Q = 17, EQ_ORDER = 3, N = 3, H = 3;
    LBM_HD static constexpr int cx(int i) { constexpr int t[Q] = {0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 2, -2, 3, -3, 0, 0}; return t[i]; }
    LBM_HD static constexpr int cy(int i) { constexpr int t[Q] = {0, 0, 0, 1, -1, 1, -1, -1, 1, 2, -2, -2, 2, 0, 0, 3, -3}; return t[i]; }
    LBM_HD static constexpr int opp(int i) { return opp_generic(i); }
    static constexpr bool UNIT_PRESSURE = false;
};
template <>
struct Lat<LBM_D2Q21> {  // src/quadratures/D2Q21.jl:15-71 (25 stored populations, last 4 weight 0)
    static constexpr int Q = 25, EQ_ORDER = 3, N = 3, H = 3;
    LBM_HD static constexpr int cx(int i) {
        constexpr int t[Q] = {0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, 2, -2, 3, -3, 0, 0, 3, -3, -3, 3};
        return t[i];
    }
    LBM_HD static constexpr int cy(int i) {
        constexpr int t[Q] = {0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 2, -2, -2, 2, 0, 0, 3, -3, 3, -3, 3, -3};
        return t[i];
    }
    LBM_HD static constexpr int opp(int i) { return opp_generic(i); }
    static constexpr bool UNIT_PRESSURE = false;
};
template <>
struct Lat<LBM_D2Q37> {  // src/quadratures/D2Q37.jl:11-74
    static constexpr int Q = 37, EQ_ORDER = 4, N = 4, H = 3;
    LBM_HD static constexpr int cx(int i) {
        constexpr int t[Q] = {0, 1, -1, 0, 0, 1, -1, 1, -1, 2, -2, 0, 0, 2, -2, -2, 2, 1, -1, 1, -1,
                              2, -2, 2, -2, 3, -3, 0, 0, 3, -3, 3, -3, 1, -1, -1, 1};
        return t[i];
    }
    LBM_HD static constexpr int cy(int i) {
        constexpr int t[Q] = {0, 0, 0, 1, -1, 1, -1, -1, 1, 0, 0, 2, -2, 1, -1, 1, -1, 2, -2, -2, 2,
                              2, -2, -2, 2, 0, 0, 3, -3, 1, -1, -1, 1, 3, -3, 3, -3};
        return t[i];
    }
    LBM_HD static constexpr int opp(int i) { return opp_generic(i); }
    static constexpr bool UNIT_PRESSURE = false;
};

// Runtime view of one lattice (host side).
struct LatticeInfo {
    int id, Q, eq_order, N, H;
    int cx[LBM_MAX_Q], cy[LBM_MAX_Q], opp[LBM_MAX_Q];
    double w[LBM_MAX_Q];
    double css;
    const char *name;
};

template <int ID>
inline void fill_structure(LatticeInfo &li) {
    using L = Lat<ID>;
    li.id = ID; li.Q = L::Q; li.eq_order = L::EQ_ORDER; li.N = L::N; li.H = L::H;
    for (int i = 0; i < L::Q; ++i) { li.cx[i] = L::cx(i); li.cy[i] = L::cy(i); li.opp[i] = L::opp(i); }
}

inline void rep(double *w, int &k, double v, int n) { for (int j = 0; j < n; ++j) w[k++] = v; }

// weights / speed_of_sound_squared with the reference's arithmetic (file:line above).
inline bool lattice_info(int id, LatticeInfo &li) {
    int k = 0;
    double *w = li.w;
    for (int i = 0; i < LBM_MAX_Q; ++i) { w[i] = 0; li.cx[i] = li.cy[i] = li.opp[i] = 0; }
    switch (id) {
    case LBM_D2Q4: fill_structure<LBM_D2Q4>(li); li.name = "D2Q4"; rep(w, k, 1.0 / 4, 4); li.css = 2.0; break;
    case LBM_D2Q5: fill_structure<LBM_D2Q5>(li); li.name = "D2Q5"; rep(w, k, 4.0 / 6, 1); rep(w, k, 1.0 / 12, 4); li.css = 6.0; break;
    case LBM_D2Q9:
        fill_structure<LBM_D2Q9>(li); li.name = "D2Q9";
        w[0] = 4.0 / 9;
        for (int i = 1; i < 9; ++i) w[i] = (i % 2) ? 1.0 / 36 : 1.0 / 9;
        li.css = 3.0;
        break;
    case LBM_D2Q13:
        fill_structure<LBM_D2Q13>(li); li.name = "D2Q13";
        rep(w, k, 3.0 / 8, 1); rep(w, k, 1.0 / 12, 4); rep(w, k, 1.0 / 16, 4); rep(w, k, 1.0 / 96, 4);
        li.css = 2.0;
        break;
    case LBM_D2Q17: {
        fill_structure<LBM_D2Q17>(li); li.name = "D2Q17";
        const double sq = std::sqrt(193.0);
        rep(w, k, (575 + 193 * sq) / 8100, 1); rep(w, k, (3355 - 91 * sq) / 18000, 4);
        rep(w, k, (655 + 17 * sq) / 27000, 4); rep(w, k, (685 - 49 * sq) / 54000, 4);
        rep(w, k, (1445 - 101 * sq) / 162000, 4);
        li.css = (125 + 5 * std::sqrt(193.0)) / 72;
        break;
    }
    case LBM_D2Q21:
        fill_structure<LBM_D2Q21>(li); li.name = "D2Q21";
        rep(w, k, 91.0 / 324, 1); rep(w, k, 1.0 / 12, 4); rep(w, k, 2.0 / 27, 4); rep(w, k, 7.0 / 360, 4);
        rep(w, k, 1.0 / 432, 4); rep(w, k, 1.0 / 1620, 4); rep(w, k, 0.0, 4);
        li.css = 3.0 / 2;
        break;
    case LBM_D2Q37: {
        fill_structure<LBM_D2Q37>(li); li.name = "D2Q37";
        rep(w, k, 0.23315066913235250228650, 1); rep(w, k, 0.10730609154221900241246, 4);
        rep(w, k, 0.05766785988879488203006, 4); rep(w, k, 0.01420821615845075026469, 4);
        rep(w, k, 0.00535304900051377523273, 8); rep(w, k, 0.00101193759267357547541, 4);
        rep(w, k, 0.00024530102775771734547, 4); rep(w, k, 0.00028341425299419821740, 8);
        const double r = 1.19697977039307435897239;
        li.css = r * r;
        break;
    }
    default: return false;
    }
    return true;
}

}  // namespace lbm
