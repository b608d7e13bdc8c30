// pb_span.h - the arithmetic core of the bit-exact ordered sums (pb_ordered.cu), host + device.
//
// A left-to-right f64 accumulation s_i = fl(s_{i-1} + a_i) (matrix2D.c:222-228, pca.c:88-93,
// cluster.c:135-148 in the reference) is modelled on integers.  Fix a unit u = 2^(eref-52), the ulp of
// the LOWEST binade the running sum visits inside a run of elements, and write s = S * u.  A step whose
// exact result x = s_{i-1} + a_i lies in binade eref + k rounds x to a multiple of g = 2^k units:
//
//     S_i = RN_g(S_{i-1} + a_i / u)                  (ties to the even multiple)
//
// S_{i-1} is a multiple of 2^k' (k' = level of the previous result).  If k <= k' it is also a multiple of
// g and the step is a TRANSLATION, S_i = S_{i-1} + RN_g(a_i / u): sequential floating-point accumulation
// degenerates into integer accumulation of quantised terms, and integer addition is associative.  Two
// things make a step depend on the state itself:
//   * a tie (the discarded part is exactly g/2): the winner is the neighbour that leaves S_i / g even;
//   * an upward step k > k': S_{i-1} has set bits below g that take part in the rounding.
// Both need ONE bit of the state when they happen on the lowest level (a tie at k = 0, a step 0 -> 1),
// and that bit is bit 0 of (S_start + everything added so far).  A run is therefore summarised for both
// parities of its start state (a two-state transducer; composition stays associative).  Anything that
// would need a higher bit (ties at k >= 1, steps from k' >= 1, steps of two or more levels) marks the run
// unusable: it is replayed sequentially.
//
// The levels k_i are PREDICTED from an approximate (unordered) running sum.  The prediction is verified,
// not trusted: every element contributes the constraint "S_start + C_i lies strictly inside binade k_i"
// (C_i = contributions so far), i.e. an interval for S_start, and so does the predicted level of the
// start state.  A run is the translation `sum` iff lo <= S_start <= hi; spans concatenate like
//     (a ++ b) = { a.sum + b.sum, max(a.lo, b.lo - a.sum), min(a.hi, b.hi - a.sum) }.
// If the exact state satisfies the interval, every rounding above used the grid the FPU uses, so the
// result is bit-identical to the sequential loop; if not, the caller falls back to the loop itself.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

#define PB_SPAN_MAX_LEVEL 8 /* levels 0..8 above the unit: |S| < 2^61 */

struct PbSpan { long long sum, lo, hi; };
struct PbSpan2 { PbSpan p[2]; }; // by bit 0 of the start state

#define PB_SPAN_INF (1LL << 61)

PB_HD long long pb_ll_max(long long a, long long b) { return a > b ? a : b; }
PB_HD long long pb_ll_min(long long a, long long b) { return a < b ? a : b; }

PB_HD PbSpan pb_span_identity() { PbSpan s; s.sum = 0; s.lo = -PB_SPAN_INF; s.hi = PB_SPAN_INF; return s; }
PB_HD PbSpan pb_span_invalid() { PbSpan s; s.sum = 0; s.lo = PB_SPAN_INF; s.hi = -PB_SPAN_INF; return s; }
PB_HD bool pb_span_valid(const PbSpan &s) { return s.lo <= s.hi; }

// a ++ b (b's constraint applies to S_start + a.sum); an empty interval is absorbing
PB_HD PbSpan pb_span_cat(const PbSpan &a, const PbSpan &b) {
    if (!pb_span_valid(a) || !pb_span_valid(b)) return pb_span_invalid();
    PbSpan r;
    r.sum = a.sum + b.sum;
    r.lo = pb_ll_max(a.lo, b.lo - a.sum);
    r.hi = pb_ll_min(a.hi, b.hi - a.sum);
    if (r.lo > r.hi) return pb_span_invalid();
    return r;
}

PB_HD PbSpan2 pb_span2_identity() { PbSpan2 r; r.p[0] = r.p[1] = pb_span_identity(); return r; }

PB_HD PbSpan2 pb_span2_cat(const PbSpan2 &a, const PbSpan2 &b) {
    PbSpan2 r;
    for (int p = 0; p < 2; p++) r.p[p] = pb_span_cat(a.p[p], b.p[(p + (int)(a.p[p].sum & 1LL)) & 1]);
    return r;
}

PB_HD long long pb_double_bits(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
PB_HD double pb_bits_double(long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
PB_HD int pb_exponent_of(double v) { return (int)((pb_double_bits(v) >> 52) & 0x7ff) - 1023; }
PB_HD double pb_pow2(int e) { return pb_bits_double((long long)(e + 1023) << 52); } // -1022 <= e <= 1023

// usable exponent range of a unit: scaling by 2^(52 - eref) must stay a normal power of two
PB_HD bool pb_eref_ok(int eref) { return eref > -900 && eref < 900; }

// a_i / u split into sign, integer part and fraction of the magnitude (all exact)
struct PbQuant {
    long long si; // signed integer part (truncated towards zero)
    double f2;    // 2 * signed fraction, |f2| < 2
    int bad;      // too large for the integer state (or NaN / Inf)
};
PB_HD PbQuant pb_quantise(double term, int eref) {
    PbQuant q;
    const double u = term * pb_pow2(52 - eref); // exact: power-of-two scaling (underflow only rounds
                                                // magnitudes below 2^-1022 units, far below any tie)
    const double au = u < 0 ? -u : u;
    q.bad = !(au < 1152921504606846976.0); // 2^60
    const long long ai = q.bad ? 0 : (long long)au;
    const double af = q.bad ? 0.0 : au - (double)ai; // exact, in [0, 1)
    q.si = u < 0 ? -ai : ai;
    q.f2 = u < 0 ? -2.0 * af : 2.0 * af;
    return q;
}

// One step on level k (g = 2^k units) from a state whose residue modulo g is Ls (0 unless this is the
// upward step 0 -> 1, where Ls = bit 0 of the state).  Returns d = S_i - S_{i-1}; *tie is set when the
// discarded part is exactly g / 2 (then d is the LOWER candidate and the caller decides).
PB_HD long long pb_round_step(const PbQuant &q, int k, long long Ls, int *tie) {
    const long long g = 1LL << k;
    const long long t = Ls + q.si;
    const long long low = t & (g - 1); // floor-mod: two's complement AND
    const double D = (double)(g - 2 * low);
    long long r;
    *tie = 0;
    if (k == 0) { // low == 0, D == 1: x = f in (-1, 1), candidates -1, 0, 1
        if (q.f2 > 1.0) r = 1;
        else if (q.f2 == 1.0) { r = 0; *tie = 1; }   // candidates 0 (lower), 1
        else if (q.f2 > -1.0) r = 0;
        else if (q.f2 == -1.0) { r = -1; *tie = 1; } // candidates -1 (lower), 0
        else r = -1;
    } else { // x = low + f in (-1, g + 1): candidates 0 and g
        if (q.f2 > D) r = g;
        else if (q.f2 == D) { r = 0; *tie = 1; }
        else r = 0;
    }
    return q.si - low + r; // S_i = (S_{i-1} - Ls) + (t - low) + r
}

#ifdef PB_SPAN_REASONS
#define PB_WHY(r, v) ((r).why = (v))
#else
#define PB_WHY(r, v) ((void)0)
#endif

// Running state of a run being summarised (one per start parity when PARITY2).
struct PbRun {
    long long C[2]; // contributions so far, by start parity
    long long lo[2], hi[2];
    int kprev;      // level of the previous result (of the start state before the first element)
    int bad;        // the run cannot be summarised
    int sensitive;  // some step depended on the parity of the state
#ifdef PB_SPAN_REASONS
    int why;        // analysis builds only: 1 level out of range, 2 term too large, 3 upward step, 4 tie
#endif
};

// constraint "S_start + C strictly inside binade level k on the side of sign(approx)"
PB_HD void pb_run_constrain(PbRun &r, int p, int k, bool neg) {
    const long long A = 1LL << (k + 52), B = 1LL << (k + 53);
    const long long lo = (neg ? -B : A) + 1 - r.C[p], hi = (neg ? -A : B) - 1 - r.C[p];
    r.lo[p] = pb_ll_max(r.lo[p], lo);
    r.hi[p] = pb_ll_min(r.hi[p], hi);
}

// start a run: approx0 = predicted state before the first element.  Every run (a thread's share of a
// block included) constrains its own start state, so neighbouring runs need not agree on the prediction:
// if they do not, the concatenated interval is empty and the block is replayed.
PB_HD void pb_run_begin(PbRun &r, double approx0, int eref) {
    r.C[0] = r.C[1] = 0;
    r.lo[0] = r.lo[1] = -PB_SPAN_INF;
    r.hi[0] = r.hi[1] = PB_SPAN_INF;
    r.bad = 0;
    r.sensitive = 0;
    const int k = pb_exponent_of(approx0) - eref;
    if (k < 0 || k > PB_SPAN_MAX_LEVEL) { r.bad = 1; r.kprev = 0; return; }
    r.kprev = k;
    pb_run_constrain(r, 0, k, approx0 < 0);
    pb_run_constrain(r, 1, k, approx0 < 0);
}

// one element; approx = predicted state AFTER it.  NV = 1: only the parity-0 variant is maintained and a
// parity-dependent step just sets `sensitive` (the run is then redone with NV = 2).
template <int NV>
PB_HD void pb_run_push(PbRun &r, double term, double approx, int eref) {
    const int k = pb_exponent_of(approx) - eref;
    if (k < 0 || k > PB_SPAN_MAX_LEVEL) { r.bad = 1; PB_WHY(r, 1); return; }
    const PbQuant q = pb_quantise(term, eref);
    if (q.bad) { r.bad = 1; PB_WHY(r, 2); return; }
    const bool up = k > r.kprev;
    if (up && !(r.kprev == 0 && k == 1)) { r.bad = 1; PB_WHY(r, 3); return; } // needs bits above bit 0 of the state
    const bool neg = approx < 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < NV; p++) {
        const long long bit0 = (p + r.C[p]) & 1LL; // bit 0 of S_{i-1}
        int tie;
        long long d = pb_round_step(q, k, up ? bit0 : 0, &tie);
        if (up) r.sensitive = 1;
        if (tie) {
            if (k != 0 || up) { r.bad = 1; PB_WHY(r, 4); return; }
            r.sensitive = 1;
            // candidates d (lower) and d + 1: the one that leaves S_i even
            if ((bit0 + d) & 1LL) d += 1;
        }
        r.C[p] += d;
        pb_run_constrain(r, p, k, neg);
    }
    r.kprev = k;
}

// The common case, cheaply: every predicted level of a thread's run (start state included) is the same
// level k and the sign never changes.  Then every step is the translation rint(a / 2^k units) and the
// constraints collapse to the extremes of the prefix sums.  Equivalent to pb_run_begin + pb_run_push<1>
// element by element (tests/native/test_span.cpp checks that); a tie marks the run sensitive (k = 0) or
// unusable (k > 0) exactly like the general path, and a sensitive run is redone by the two-parity pass.
struct PbUni {
    double ps, mn, mx; // prefix sum of the quantised terms and its extremes (0 = the start state included)
    int bad, tie;
};
PB_HD void pb_uni_begin(PbUni &u) { u.ps = u.mn = u.mx = 0.0; u.bad = u.tie = 0; }
PB_HD double pb_uni_scale(int k, int eref) { return pb_pow2(52 - eref - k); }
PB_HD void pb_uni_push(PbUni &q, double term, double scale) {
    const double MAGIC = 6755399441055744.0; // 1.5 * 2^52: (u + MAGIC) - MAGIC == rint(u) for |u| < 2^51
    const double u = term * scale;           // exact
    const double au = u < 0 ? -u : u;
    const double d = (u + MAGIC) - MAGIC;
    const double rem = u - d;
    q.bad |= !(au < 2251799813685248.0); // 2^51 (or NaN)
    q.tie |= (rem == 0.5) | (rem == -0.5);
    q.ps += d;
    q.mn = q.ps < q.mn ? q.ps : q.mn;
    q.mx = q.ps > q.mx ? q.ps : q.mx;
}
PB_HD void pb_uni_end(const PbUni &q, PbRun &r, int k, bool neg) {
    r.kprev = k;
    r.bad = q.bad | (q.tie && k != 0);
    r.sensitive = q.tie && k == 0;
    if (r.bad) { r.C[0] = r.C[1] = 0; return; }
    const long long A = 1LL << (k + 52), B = 1LL << (k + 53);
    const long long lmn = ((long long)q.mn) * (1LL << k), lmx = ((long long)q.mx) * (1LL << k);
    r.C[0] = r.C[1] = ((long long)q.ps) * (1LL << k);
    r.lo[0] = r.lo[1] = (neg ? -B : A) + 1 - lmn;
    r.hi[0] = r.hi[1] = (neg ? -A : B) - 1 - lmx;
}
PB_HD void pb_run_uniform(PbRun &r, const double *t, int cnt, int k, bool neg, int eref) {
    PbUni q;
    pb_uni_begin(q);
    const double scale = pb_uni_scale(k, eref);
    for (int i = 0; i < cnt; i++) pb_uni_push(q, t[i], scale);
    pb_uni_end(q, r, k, neg);
}

template <int NV>
PB_HD PbSpan2 pb_run_span(const PbRun &r) {
    PbSpan2 s;
    for (int p = 0; p < 2; p++) {
        const int v = NV == 2 ? p : 0;
        s.p[p].sum = r.C[v];
        s.p[p].lo = r.lo[v];
        s.p[p].hi = r.hi[v];
        if (r.bad || s.p[p].lo > s.p[p].hi) s.p[p] = pb_span_invalid();
    }
    return s;
}

// ---- block-uniform fast path (k_ord_fast in pb_ordered.cu) -------------------------------------------------
// If every partial sum of a block lies in the binade e of the block's predicted start state, every step is a
// translation on the ONE grid u = 2^(e-52), and the FPU quantises for us: RN_u(a) = (a + M) - M with
// M = 1.5 * 2^e (valid while |a| < 2^(e-1); ulp(M) = u, M / u even).  The quantised terms are multiples of u,
// so their sums are exact in floating point below 2^53 u.  A tie - |a - RN_u(a)| = u/2, its winner depends on
// the parity of the state - and anything else that is not this case leaves the pair to the general path.  As
// above the claim is an interval for the exact start state: S_start + (every in-order prefix) strictly inside
// the binade.  This is pb_run_uniform with k = 0 stretched over a whole block.
struct PbFastGrid {
    double M, half; // 1.5 * 2^e, u / 2
    int e, ok;
};
PB_HD PbFastGrid pb_fast_grid(double pstart) {
    PbFastGrid g;
    g.e = pb_exponent_of(pstart);
    g.ok = pb_eref_ok(g.e) ? 1 : 0;
    const int ee = g.ok ? g.e : 0;
    g.M = 1.5 * pb_pow2(ee);
    g.half = pb_pow2(ee - 53);
    return g;
}
// one element: returns RN_u(t); sets *tie when t sits exactly halfway between two grid points
PB_HD double pb_fast_quant(const PbFastGrid &g, double t, bool *tie) {
    const double d = (t + g.M) - g.M; // a multiple of u
    const double rem = t - d;         // exact
    const double ar = rem < 0 ? -rem : rem;
    *tie = ar == g.half;
    return d;
}
#define PB_FAST_LIMIT 1125899906842624.0 /* 2^50 units */
// sum / lo_ext / hi_ext: total and extremes (<= 0 <= ) of the in-order prefix sums of the quantised terms, in
// the input's scale.  margin: distance (units) the PREDICTED start state must keep from the interval's ends -
// a performance heuristic only (a record that will not apply costs a replay); soundness is the interval.
PB_HD bool pb_fast_finish(const PbFastGrid &g, double pstart, double sum, double lo_ext, double hi_ext, bool tie,
                          long long margin, PbSpan &out) {
    if (!g.ok || tie) return false;
    const double scale = pb_pow2(52 - g.e);
    const double su = sum * scale, lu = lo_ext * scale, hu = hi_ext * scale; // exact power-of-two scalings
    const double asu = su < 0 ? -su : su;
    // below the limit every partial sum was exact and every |a_i| < 2^(e-1) (a larger term moves a prefix by
    // at least 2^51 - 2 units); NaN fails every comparison
    if (!(asu < PB_FAST_LIMIT) || !((hu - lu) < PB_FAST_LIMIT) || !(lu <= 0.0) || !(hu >= 0.0)) return false;
    const long long A0 = 1LL << 52, B = 1LL << 53;
    const bool neg = pstart < 0.0;
    out.sum = (long long)su;
    out.lo = (neg ? -B : A0) + 1 - (long long)lu;
    out.hi = (neg ? -A0 : B) - 1 - (long long)hu;
    const long long S0 = (long long)(pstart * scale); // an integer with |S0| in [2^52, 2^53)
    return out.lo + margin <= S0 && S0 <= out.hi - margin;
}

// exact state <-> integer in units of 2^(eref-52)
PB_HD bool pb_state_to_units(double s, int eref, long long &S) {
    const long long bits = pb_double_bits(s);
    if ((bits << 1) == 0) { S = 0; return true; }
    const int es = (int)((bits >> 52) & 0x7ff) - 1023, k0 = es - eref;
    if (es < -1000 || es > 1000 || k0 < 0 || k0 > PB_SPAN_MAX_LEVEL) return false;
    const long long M = ((bits & 0x000fffffffffffffLL) | (1LL << 52)) << k0;
    S = bits < 0 ? -M : M;
    return true;
}
// valid for states that passed a span's interval check (at most 53 significant bits)
PB_HD double pb_units_to_state(long long S, int eref) { return (double)S * pb_pow2(eref - 52); }

// apply a summarised run to the exact state; false = not applicable (replay it)
PB_HD bool pb_span2_apply(const PbSpan2 &sp, int eref, double &s) {
    long long S;
    if (!pb_eref_ok(eref) || !pb_state_to_units(s, eref, S)) return false;
    const PbSpan &v = sp.p[(int)(S & 1LL)];
    if (!(S >= v.lo && S <= v.hi)) return false;
    s = pb_units_to_state(S + v.sum, eref);
    return true;
}

// ---- the resolving walker's state: an exact double kept as integer * unit ---------------------------
// value = S * 2^(e - 52).  Between records of the same unit the walk is pure integer arithmetic; the
// double is only rebuilt where the unit changes or a block has to be replayed.
struct PbState {
    long long S;
    int e;
    int ok; // 0: the value is zero, subnormal, huge or non-finite - records cannot be applied to it
};
PB_HD PbState pb_state_from_double(double s) {
    PbState st;
    const long long bits = pb_double_bits(s);
    const int ef = (int)((bits >> 52) & 0x7ff);
    st.e = ef - 1023;
    st.ok = ef > 123 && ef < 1923; // |e| < 900
    const long long M = (bits & 0x000fffffffffffffLL) | (1LL << 52);
    st.S = bits < 0 ? -M : M;
    return st;
}
PB_HD double pb_state_to_double(const PbState &st) { return (double)st.S * pb_pow2(st.e - 52); }
// re-express in unit eref (exact) - false if the value has bits below it or would overflow the level range.
// Integer-only (this sits on the resolving warp's critical path): the value's binade is e - 52 + (index of
// the top set bit of |S|); a valid state has at most 53 significant bits, so the shift is exact whenever
// that binade is not below eref.
PB_HD bool pb_state_rebase(PbState &st, int eref) {
    if (!st.ok) return false;
    if (st.e == eref) return true;
    if (!pb_eref_ok(eref)) return false;
    if (st.S == 0) { st.e = eref; return true; }
    const unsigned long long a = (unsigned long long)(st.S < 0 ? -st.S : st.S);
#if defined(__CUDA_ARCH__)
    const int msb = 63 - __clzll((long long)a);
#else
    const int msb = 63 - __builtin_clzll(a);
#endif
    const int es = st.e - 52 + msb, k0 = es - eref;
    if (es < -1000 || es > 1000 || k0 < 0 || k0 > PB_SPAN_MAX_LEVEL) return false;
    const int d = st.e - eref; // k0 - (msb - 52): the result has msb + d = k0 + 52 <= 60
    const unsigned long long r = d >= 0 ? a << d : a >> -d;
    st.S = st.S < 0 ? -(long long)r : (long long)r;
    st.e = eref;
    return true;
}
// the previous formulation, through the double (kept for the self-test)
PB_HD bool pb_state_rebase_ref(PbState &st, int eref) {
    if (!st.ok) return false;
    if (st.e == eref) return true;
    long long S;
    const double v = pb_state_to_double(st);
    if (!pb_eref_ok(eref) || !pb_state_to_units(v, eref, S)) return false;
    st.S = S;
    st.e = eref;
    return true;
}
// one record (both parities) applied to the state; false = replay the block
PB_HD bool pb_state_apply(PbState &st, const PbSpan &v0, const PbSpan &v1, int eref) {
    if (!pb_state_rebase(st, eref)) return false;
    const PbSpan &v = (st.S & 1LL) ? v1 : v0;
    if (!(st.S >= v.lo && st.S <= v.hi)) return false; // an invalid span has lo > hi
    st.S += v.sum;
    return true;
}
