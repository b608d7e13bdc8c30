// CPU check of patolette_b200/csrc/pb_span.h: the block-summary arithmetic of the ordered sums,
// emulated with the kernel's tiling (blocks of 512 elements, 16 consecutive elements per lane,
// in-order tree composition), against the literal sequential loop.
//
//   g++ -O2 -ffp-contract=off -std=c++17 -I patolette_b200/csrc tests/native/test_span.cpp -o /tmp/test_span
//
// For every chain and block: if the summary accepts the exact start state, the state it produces must
// be bit-identical to the sequential loop's.  Prints acceptance statistics per data family; exits 1 on
// any mismatch.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <string>
#include <vector>

#include "pb_span.h"

static const int OB = 512, PER = 16, THREADS = OB / PER;

struct Stats {
    long blocks = 0, accepted = 0, wrong = 0, sensitive = 0, bad = 0, assoc_fail = 0, fast = 0, fast_mismatch = 0;
};

static double tree_sum(const double *a, int n) {
    if (n <= 0) return 0.0;
    if (n == 1) return a[0];
    return tree_sum(a, n / 2) + tree_sum(a + n / 2, n - n / 2);
}

static void summarise_block(const double *a, int cnt, double pstart, PbSpan2 &out, int &eref, bool &usable,
                            bool &sensitive, Stats &st) {
    // approximate running sums exactly as the kernel predicts them
    double tl[THREADS], tstart[THREADS];
    for (int t = 0; t < THREADS; t++) {
        double s = 0;
        for (int k = 0; k < PER; k++)
            if (t * PER + k < cnt) s += a[t * PER + k];
        tl[t] = s;
    }
    {
        double run = 0;
        for (int t = 0; t < THREADS; t++) { tstart[t] = pstart + run; run += tl[t]; }
    }
    int emin = 1 << 20, emax = -(1 << 20);
    for (int t = 0; t < THREADS; t++) {
        if (t * PER >= cnt) break;
        double run = tstart[t];
        int e = pb_exponent_of(run);
        emin = e < emin ? e : emin; emax = e > emax ? e : emax;
        for (int k = 0; k < PER; k++)
            if (t * PER + k < cnt) {
                run += a[t * PER + k];
                e = pb_exponent_of(run);
                emin = e < emin ? e : emin; emax = e > emax ? e : emax;
            }
    }
    eref = emin;
    usable = pb_eref_ok(emin) && pb_eref_ok(emax) && emax - emin <= PB_SPAN_MAX_LEVEL;
    sensitive = false;
    out = pb_span2_identity();
    if (!usable) { st.bad++; return; }
    std::vector<PbSpan2> spans;
    for (int pass = 0; pass < 2; pass++) { // pass 0: single variant; pass 1: both parities if sensitive
        spans.clear();
        bool sens = false, bad = false;
        for (int t = 0; t < THREADS; t++) {
            if (t * PER >= cnt) break;
            PbRun r;
            pb_run_begin(r, tstart[t], eref);
            double run = tstart[t];
            int lmin = pb_exponent_of(run), lmax = lmin, mycnt = 0;
            bool sign_same = true;
            const bool neg0 = run < 0;
            for (int k = 0; k < PER; k++)
                if (t * PER + k < cnt) {
                    run += a[t * PER + k];
                    mycnt++;
                    const int e = pb_exponent_of(run);
                    lmin = e < lmin ? e : lmin; lmax = e > lmax ? e : lmax;
                    sign_same &= (run < 0) == neg0;
                    if (pass == 0) pb_run_push<1>(r, a[t * PER + k], run, eref);
                    else pb_run_push<2>(r, a[t * PER + k], run, eref);
                }
            if (pass == 0 && lmin == lmax && sign_same) { // the kernel takes the cheap path here: must agree
                PbRun f;
                pb_run_uniform(f, &a[t * PER], mycnt, lmin - eref, neg0, eref);
                st.fast++;
                const bool same = f.bad ? true /* stricter magnitude limit is allowed */
                                        : (!r.bad && f.sensitive == r.sensitive &&
                                           (f.sensitive || (f.C[0] == r.C[0] && f.lo[0] == r.lo[0] && f.hi[0] == r.hi[0])));
                if (!same) {
                    st.fast_mismatch++;
                    if (st.fast_mismatch < 4)
                        fprintf(stderr, "  FAST/SLOW mismatch: bad %d/%d sens %d/%d C %lld/%lld lo %lld/%lld hi %lld/%lld\n", f.bad, r.bad,
                                f.sensitive, r.sensitive, f.C[0], r.C[0], f.lo[0], r.lo[0], f.hi[0], r.hi[0]);
                }
                r = f; // the kernel uses the cheap path's result
            }
            sens |= r.sensitive != 0;
            bad |= r.bad != 0;
            spans.push_back(pass == 0 ? pb_run_span<1>(r) : pb_run_span<2>(r));
        }
        if (bad) { usable = false; st.bad++; return; }
        if (pass == 0 && !sens) break;
        if (pass == 0) continue; // parity-dependent: every thread's span is redone for both of ITS start parities
        // a block whose start state sits above its lowest binade (enforced by its start constraint) only
        // ever sees block-level parity 0: the record can be stored as a plain one (variant 0 of the composition)
        sensitive = pb_exponent_of(pstart) - eref < 1;
    }
    // in-order composition: left fold and a balanced tree must agree (associativity)
    PbSpan2 fold = spans[0];
    for (size_t i = 1; i < spans.size(); i++) fold = pb_span2_cat(fold, spans[i]);
    std::vector<PbSpan2> lvl = spans;
    while (lvl.size() > 1) {
        std::vector<PbSpan2> nx;
        for (size_t i = 0; i < lvl.size(); i += 2)
            nx.push_back(i + 1 < lvl.size() ? pb_span2_cat(lvl[i], lvl[i + 1]) : lvl[i]);
        lvl.swap(nx);
    }
    for (int p = 0; p < 2; p++) {
        const PbSpan &x = fold.p[p], &y = lvl[0].p[p];
        const bool vx = pb_span_valid(x), vy = pb_span_valid(y);
        if (vx != vy || (vx && (x.sum != y.sum || x.lo != y.lo || x.hi != y.hi))) st.assoc_fail++;
    }
    out = fold;
    if (sensitive) st.sensitive++;
}

static void run_chain(const std::vector<double> &a, Stats &st) {
    const size_t n = a.size();
    const size_t nblk = (n + OB - 1) / OB;
    std::vector<double> bsum(nblk), pstart(nblk);
    for (size_t b = 0; b < nblk; b++) bsum[b] = tree_sum(&a[b * OB], (int)std::min<size_t>(OB, n - b * OB));
    {
        double run = 0;
        for (size_t b = 0; b < nblk; b++) { pstart[b] = run; run += bsum[b]; }
    }
    // the resolving walk as the kernel does it: integer state, rebuilt from the double only after a replay
    double s = 0.0;
    PbState state = pb_state_from_double(s);
    for (size_t b = 0; b < nblk; b++) {
        const int cnt = (int)std::min<size_t>(OB, n - b * OB);
        double truth = s;
        for (int i = 0; i < cnt; i++) truth = truth + a[b * OB + i];
        PbSpan2 sp;
        int eref;
        bool usable, sens;
        summarise_block(&a[b * OB], cnt, pstart[b], sp, eref, usable, sens, st);
        st.blocks++;
        bool applied = false;
        if (usable) {
            // a parity-dependent block whose start state sits above its lowest binade only ever sees parity 0
            const bool plain = !sens;
            applied = pb_state_apply(state, sp.p[0], plain ? sp.p[0] : sp.p[1], eref);
            if (applied) {
                st.accepted++;
                const double got = pb_state_to_double(state);
                if (pb_double_bits(got) != pb_double_bits(truth)) {
                    st.wrong++;
                    if (st.wrong < 5)
                        fprintf(stderr, "  MISMATCH block %zu: start %a truth %a got %a eref %d sens %d\n", b, s, truth, got,
                                eref, (int)sens);
                    state = pb_state_from_double(truth);
                }
            }
        }
        if (!applied) state = pb_state_from_double(truth); // replay = the literal loop
        s = truth;
    }
}

// The run records of k_ord_group / k_ord_resolve (pb_ordered.cu), emulated lane by lane: a backward
// segmented doubling scan composes, for every record of a group of 32, the maximal run of usable records of
// one unit that starts there (both start parities); the walk crosses a whole run with one interval check on
// the variant its state's parity selects.  Whatever it accepts must be the sequential loop's state.
struct RunStats { long runs = 0, run_blocks = 0, singles = 0, replays = 0, wrong = 0; };
static void run_chain_with_run_records(const std::vector<double> &a, RunStats &rs) {
    const size_t n = a.size(), nblk = (n + OB - 1) / OB;
    std::vector<double> pstart(nblk), truth(nblk);
    {
        double run = 0, s = 0;
        for (size_t b = 0; b < nblk; b++) {
            const int cnt = (int)std::min<size_t>(OB, n - b * OB);
            pstart[b] = run;
            run += tree_sum(&a[b * OB], cnt);
            for (int i = 0; i < cnt; i++) s = s + a[b * OB + i];
            truth[b] = s;
        }
    }
    enum { OK = 0, SENS = 1, REPLAY = 2 };
    std::vector<PbSpan2> rec(nblk);
    std::vector<int> eref(nblk), kind(nblk);
    Stats dummy;
    for (size_t b = 0; b < nblk; b++) {
        const int cnt = (int)std::min<size_t>(OB, n - b * OB);
        bool usable, sens;
        summarise_block(&a[b * OB], cnt, pstart[b], rec[b], eref[b], usable, sens, dummy);
        kind[b] = !usable ? REPLAY : (sens ? SENS : OK);
        if (kind[b] == OK) { rec[b].p[1] = rec[b].p[0]; if (!pb_span_valid(rec[b].p[0])) kind[b] = REPLAY; }
        if (kind[b] == SENS && !pb_span_valid(rec[b].p[0]) && !pb_span_valid(rec[b].p[1])) kind[b] = REPLAY;
    }
    PbState st = pb_state_from_double(0.0);
    auto check = [&](size_t b) {
        if (pb_double_bits(pb_state_to_double(st)) != pb_double_bits(truth[b])) { rs.wrong++; st = pb_state_from_double(truth[b]); }
    };
    for (size_t g0 = 0; g0 < nblk; g0 += 32) {
        const int gcnt = (int)std::min<size_t>(32, nblk - g0);
        PbSpan2 v[32], nv[32];
        int len[32], nlen[32], er[32];
        bool open[32], nopen[32];
        for (int l = 0; l < 32; l++) {
            v[l] = pb_span2_identity(); len[l] = 0; er[l] = 0;
            if (l < gcnt) { v[l] = rec[g0 + l]; er[l] = eref[g0 + l]; len[l] = kind[g0 + l] != REPLAY ? 1 : 0; }
            open[l] = len[l] > 0;
        }
        for (int o = 1; o < 32; o <<= 1) { // every lane reads its partner's values of the previous round
            for (int l = 0; l < 32; l++) {
                nv[l] = v[l]; nlen[l] = len[l]; nopen[l] = open[l];
                if (!open[l]) continue;
                const int p = l + o;
                if (p < 32 && len[p] > 0 && er[p] == er[l]) { nv[l] = pb_span2_cat(v[l], v[p]); nlen[l] = len[l] + len[p]; nopen[l] = open[p]; }
                else nopen[l] = false;
            }
            for (int l = 0; l < 32; l++) { v[l] = nv[l]; len[l] = nlen[l]; open[l] = nopen[l]; }
        }
        int next = 0;
        while (next < gcnt) {
            if (len[next] > 0 && pb_state_rebase(st, er[next])) {
                const PbSpan &pick = v[next].p[(int)(st.S & 1LL)];
                if (st.S >= pick.lo && st.S <= pick.hi) {
                    st.S += pick.sum;
                    rs.runs++; rs.run_blocks += len[next];
                    next += len[next];
                    check(g0 + next - 1);
                    continue;
                }
            }
            const size_t b = g0 + next;
            bool applied = false;
            if (kind[b] != REPLAY) applied = pb_state_apply(st, rec[b].p[0], rec[b].p[1], eref[b]);
            if (applied) { rs.singles++; check(b); }
            else { rs.replays++; st = pb_state_from_double(truth[b]); }
            next++;
        }
    }
}

// k_ord_fast (pb_ordered.cu) emulated lane by lane: block-uniform quantisation on the grid of the predicted
// start state, in-order prefixes inside a lane, span-style composition across lanes, pb_fast_finish.  Whatever
// it accepts AND the exact state satisfies must be the sequential loop's result.
struct FastStats { long blocks = 0, offered = 0, applied = 0, refused_by_state = 0, wrong = 0, mono_blocks = 0; };
static void run_chain_fast(const std::vector<double> &a, FastStats &fs) {
    const size_t n = a.size(), nblk = (n + OB - 1) / OB;
    bool all_nonneg = true;
    for (double v : a) all_nonneg &= v >= 0;
    double run = 0, s = 0;
    for (size_t b = 0; b < nblk; b++) {
        const int cnt = (int)std::min<size_t>(OB, n - b * OB);
        const double pstart = run;
        run += tree_sum(&a[b * OB], cnt);
        double truth = s;
        for (int i = 0; i < cnt; i++) truth = truth + a[b * OB + i];
        fs.blocks++;
        const PbFastGrid g = pb_fast_grid(pstart);
        double ps[THREADS], mn[THREADS], mx[THREADS];
        bool tie = false;
        for (int t = 0; t < THREADS; t++) {
            ps[t] = mn[t] = mx[t] = 0.0;
            for (int k = 0; k < PER; k++)
                if (t * PER + k < cnt) {
                    bool tc;
                    ps[t] += pb_fast_quant(g, a[b * OB + t * PER + k], &tc);
                    tie |= tc;
                    mn[t] = ps[t] < mn[t] ? ps[t] : mn[t];
                    mx[t] = ps[t] > mx[t] ? ps[t] : mx[t];
                }
        }
        double tot = 0, lo = 0, hi = 0;
        if (all_nonneg && (b & 1)) { // the kernel's shortcut for chains whose terms are never negative
            for (int t = 0; t < THREADS; t++) tot += ps[t];
            lo = 0; hi = tot;
            fs.mono_blocks++;
        } else {
            double excl = 0;
            for (int t = 0; t < THREADS; t++) {
                lo = excl + mn[t] < lo ? excl + mn[t] : lo;
                hi = excl + mx[t] > hi ? excl + mx[t] : hi;
                excl += ps[t];
            }
            tot = excl;
        }
        PbSpan sp;
        if (pb_fast_finish(g, pstart, tot, lo, hi, tie, 1LL << 20, sp)) {
            fs.offered++;
            PbState st = pb_state_from_double(s);
            if (pb_state_apply(st, sp, sp, g.e)) {
                fs.applied++;
                if (pb_double_bits(pb_state_to_double(st)) != pb_double_bits(truth)) {
                    fs.wrong++;
                    if (fs.wrong < 5) fprintf(stderr, "  FAST MISMATCH block %zu: start %a truth %a got %a\n", b, s, truth, pb_state_to_double(st));
                }
            } else fs.refused_by_state++;
        }
        s = truth;
    }
}

// pb_state_rebase (integer shifts) must agree with the formulation through the double for every valid
// state (at most 53 significant bits, any trailing-zero count up to the level range) and every unit.
static long check_rebase(std::mt19937_64 &rng) {
    long wrong = 0;
    for (int it = 0; it < 4000000; it++) {
        PbState a;
        const int tz = (int)(rng() % 10), bits = 1 + (int)(rng() % 53);
        unsigned long long m = rng() >> (64 - bits);
        m |= 1ULL << (bits - 1);
        if (bits + tz > 61) continue;
        a.S = (long long)(m << tz);
        if (rng() & 1) a.S = -a.S;
        if (rng() % 64 == 0) a.S = 0;
        a.e = (int)(rng() % 1900) - 950;
        a.ok = (rng() % 32) != 0;
        const int eref = a.e + (int)(rng() % 25) - 12 + ((rng() % 16 == 0) ? 900 : 0);
        PbState b = a;
        const bool ra = pb_state_rebase(a, eref), rb = pb_state_rebase_ref(b, eref);
        if (ra != rb || (ra && (a.S != b.S || a.e != b.e))) wrong++;
    }
    printf("rebase self-test: %ld mismatches\n", wrong);
    return wrong;
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 6;
    std::mt19937_64 rng(12345);
    if (check_rebase(rng)) { printf("FAILED\n"); return 1; }
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> G(0.0, 1.0);
    struct Family { std::string name; int kind; };
    const Family fams[] = {
        {"uniform positive", 0},      {"weighted positive (w in 1..1025)", 1}, {"centred products (hovering)", 2},
        {"centred squares", 3},       {"dyadic 2^-8 (ties)", 4},               {"dyadic products hovering", 5},
        {"tiny cancellation", 6},     {"sign-changing drift", 7},              {"spikes", 8},
        {"sorted ascending", 9},      {"integers (exact)", 10},                {"wide dynamic range", 11},
        {"alternating near power of two", 12},     {"mean-centred products (zero-drift walk)", 13},
        {"mean-centred, weakly correlated", 14},
    };
    long total_wrong = 0, total_assoc = 0;
    RunStats rr;
    long fast_wrong = 0;
    for (const Family &f : fams) {
        Stats st;
        FastStats fs;
        for (int rep = 0; rep < reps; rep++) {
            const size_t n = (size_t)(1000 + (rng() % 400000));
            std::vector<double> a(n);
            const double mu = U(rng), mu2 = U(rng);
            for (size_t i = 0; i < n; i++) {
                double v = 0;
                switch (f.kind) {
                case 0: v = U(rng); break;
                case 1: v = U(rng) * (double)(1 + rng() % 1025); break;
                case 2: v = (U(rng) - mu) * (U(rng) - mu2); break;
                case 3: { double d = U(rng) - mu; v = d * d; break; }
                case 4: v = (double)(rng() % 257) / 256.0; break;
                case 5: v = ((double)(rng() % 257) / 256.0 - 0.5) * ((double)(rng() % 257) / 256.0 - 0.5); break;
                case 6: v = (i & 1) ? 1.0 + 1e-9 * U(rng) : -1.0 + 1e-9 * U(rng); break;
                case 7: v = G(rng) + 0.01 * sin((double)i * 1e-4); break;
                case 8: v = (rng() % 5000 == 0) ? 1e6 * U(rng) : U(rng) * 1e-3; break;
                case 9: v = (double)i / (double)n; break;
                case 10: v = (double)(long)(rng() % 2001) - 1000.0; break;
                case 11: v = ldexp(G(rng), (int)(rng() % 60) - 30); break;
                case 12: v = ((i & 1) ? -1.0 : 1.0) * (0.25 + 1e-3 * U(rng)) + ((i % 64 == 0) ? 1.0 / 64 : 0.0); break;
                }
                a[i] = v;
            }
            if (f.kind == 13 || f.kind == 14) { // off-diagonal covariance terms of (nearly) independent channels
                std::vector<double> x(n), y(n);
                double mx = 0, my = 0;
                for (size_t i = 0; i < n; i++) {
                    x[i] = U(rng);
                    y[i] = f.kind == 14 ? 0.02 * x[i] + U(rng) : U(rng);
                    mx += x[i]; my += y[i];
                }
                mx *= 1.0 / (double)n; my *= 1.0 / (double)n;
                for (size_t i = 0; i < n; i++) a[i] = (x[i] - mx) * (y[i] - my);
            }
            if (f.kind == 12) a[0] = 1.0; // start right at a power of two and wobble around it
            run_chain(a, st);
            run_chain_with_run_records(a, rr);
            run_chain_fast(a, fs);
            if (rep == 0) { // the same data far along a chain (start state 2^20 times the terms) and behind a negative state
                std::vector<double> b2(a);
                b2.insert(b2.begin(), (f.kind % 2 ? -1.0 : 1.0) * 1048576.0 * (1.0 + U(rng)));
                run_chain_fast(b2, fs);
            }
        }
        printf("    fast path: offered %6.2f%% of blocks, applied %6.2f%%, refused by the exact state %ld, wrong %ld\n",
               100.0 * fs.offered / fs.blocks, 100.0 * fs.applied / fs.blocks, fs.refused_by_state, fs.wrong);
        fast_wrong += fs.wrong;
        printf("%-40s blocks %7ld accepted %6.2f%% sensitive %5.2f%% unusable %5.2f%% fast-threads %5.1f%% wrong %ld assoc %ld fastmis %ld\n",
               f.name.c_str(), st.blocks, 100.0 * st.accepted / st.blocks, 100.0 * st.sensitive / st.blocks,
               100.0 * st.bad / st.blocks, 100.0 * st.fast / (st.blocks * (double)THREADS), st.wrong, st.assoc_fail, st.fast_mismatch);
        total_wrong += st.wrong + st.fast_mismatch;
        total_assoc += st.assoc_fail;
    }
    printf("run records: %ld runs crossing %ld blocks, %ld single records, %ld replays, wrong %ld\n", rr.runs, rr.run_blocks,
           rr.singles, rr.replays, rr.wrong);
    total_wrong += rr.wrong + fast_wrong;
    if (total_wrong || total_assoc) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
