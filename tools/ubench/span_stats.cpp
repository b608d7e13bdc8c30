// Offline analysis of the ordered-sum block summaries on real chains (tools/analyse_chains.py):
// how many blocks are usable / parity-dependent / unusable, and why, for a given block size.
//   g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC -DPB_SPAN_REASONS -I patolette_b200/csrc tools/ubench/span_stats.cpp -o /tmp/libspan_stats.so
#include <algorithm>
#include <vector>
#include "pb_span.h"

static double tree_sum(const double *a, int n) {
    if (n <= 0) return 0.0;
    if (n == 1) return a[0];
    return tree_sum(a, n / 2) + tree_sum(a + n / 2, n - n / 2);
}

// out: 0 blocks, 1 accepted, 2 wrong, 3 unusable: range/spread, 4 unusable: term too large, 5 unusable: upward step,
//      6 unusable: tie above level 0, 7 sensitive (two-parity record), 8 plainified, 9 unit changes, 10 interval failures,
//      11 first-block, 12 elements in replayed blocks,
//      13 groups of 32 records, 14 groups that are all plain records of one unit (pass the level-2 scan today),
//      15 groups whose records are all usable (plain or two-parity) with one unit, 16 the same ignoring the unit,
//      17 level-1 runs (maximal runs of usable records of one unit inside a group), 18 records in those runs
extern "C" void span_stats(const double *a, long n, int OB, int PER, long *out) {
    const long nblk = (n + OB - 1) / OB;
    std::vector<double> pstart(nblk);
    {
        double run = 0;
        for (long b = 0; b < nblk; b++) { pstart[b] = run; run += tree_sum(a + b * OB, (int)std::min<long>(OB, n - b * OB)); }
    }
    double s = 0.0;
    PbState state = pb_state_from_double(s);
    int prev_e = 1 << 30;
    std::vector<int> kind(nblk, 2), unit(nblk, 0); // 0 plain, 1 two-parity, 2 unusable
    for (long b = 0; b < nblk; b++) {
        const int cnt = (int)std::min<long>(OB, n - b * OB);
        const double *x = a + b * OB;
        double truth = s;
        for (int i = 0; i < cnt; i++) truth += x[i];
        const int T = (cnt + PER - 1) / PER;
        std::vector<double> ts(T);
        {
            double r = 0;
            for (int t = 0; t < T; t++) {
                ts[t] = pstart[b] + r;
                double q = 0;
                for (int k = 0; k < PER && t * PER + k < cnt; k++) q += x[t * PER + k];
                r += q;
            }
        }
        int emin = 1 << 20, emax = -(1 << 20);
        for (int t = 0; t < T; t++) {
            double r = ts[t];
            int e = pb_exponent_of(r);
            emin = std::min(emin, e); emax = std::max(emax, e);
            for (int k = 0; k < PER && t * PER + k < cnt; k++) {
                r += x[t * PER + k];
                e = pb_exponent_of(r);
                emin = std::min(emin, e); emax = std::max(emax, e);
            }
        }
        out[0]++;
        bool applied = false;
        const bool usable = pb_eref_ok(emin) && pb_eref_ok(emax) && emax - emin <= PB_SPAN_MAX_LEVEL;
        if (!usable) {
            out[3]++;
            if (b == 0) out[11]++;
        } else {
            const int eref = emin;
            PbSpan2 fold = pb_span2_identity();
            bool sens = false;
            int why = 0;
            for (int t = 0; t < T; t++) {
                PbRun r;
                r.why = 0;
                pb_run_begin(r, ts[t], eref);
                if (r.bad && !why) why = 1;
                double rr = ts[t];
                for (int k = 0; k < PER && t * PER + k < cnt; k++) {
                    rr += x[t * PER + k];
                    if (!r.bad) pb_run_push<2>(r, x[t * PER + k], rr, eref);
                }
                if (r.bad && !why) why = r.why ? r.why : 1;
                sens |= r.sensitive != 0;
                fold = pb_span2_cat(fold, pb_run_span<2>(r));
            }
            if (why) {
                out[why == 1 ? 3 : why == 2 ? 4 : why == 3 ? 5 : 6]++;
            } else {
                const bool low = pb_exponent_of(pstart[b]) - eref < 1;
                kind[b] = (sens && low) ? 1 : 0;
                unit[b] = eref;
                if (sens && low) out[7]++;
                if (sens && !low) out[8]++;
                if (prev_e != (1 << 30) && prev_e != eref) out[9]++;
                prev_e = eref;
                applied = pb_state_apply(state, fold.p[0], (sens && low) ? fold.p[1] : fold.p[0], eref);
                if (applied) {
                    out[1]++;
                    if (pb_double_bits(pb_state_to_double(state)) != pb_double_bits(truth)) out[2]++;
                } else {
                    out[10]++;
                }
            }
        }
        if (!applied) { state = pb_state_from_double(truth); out[12] += cnt; }
        s = truth;
    }
    for (long g0 = 0; g0 < nblk; g0 += 32) {
        const long g1 = std::min<long>(g0 + 32, nblk);
        bool plain = true, usable = true, one_unit = true;
        for (long b = g0; b < g1; b++) {
            plain &= kind[b] == 0;
            usable &= kind[b] != 2;
            one_unit &= unit[b] == unit[g0];
        }
        out[13]++;
        out[14] += plain && one_unit;
        out[15] += usable && one_unit;
        out[16] += usable;
        for (long b = g0; b < g1;) {
            if (kind[b] == 2) { b++; continue; }
            long e = b + 1;
            while (e < g1 && kind[e] != 2 && unit[e] == unit[b]) e++;
            out[17]++;
            out[18] += e - b;
            b = e;
        }
    }
}
