// lgar_c_b200/csrc/lgar_search.cuh -- the two psi searches of the reference, and how the engine gets through them fast.
//
// lgar_theta_mass_balance (reference src/lgar.cxx:2952-3112) and lgar_theta_mass_balance_correction (:3262-3435) look for the
// capillary head psi at which the water held by a set of fronts equals a prior mass, by a decimal DIGIT search: psi moves in
// steps of 0.1*factor, the step shrinks by 10 whenever the search turns from "too wet" to "too dry", and it stops when the mass
// is within 1e-10 cm (or on one of six other conditions).  The result is whatever psi the walk stops at, so it is defined by the
// walk itself and has to be reproduced trip for trip.  On the bench workload EVERY iterated search takes
// 32-128 trips (mean 85, tail to 10^4; tools/search_trip_histogram.py) and 90 % of all theta(psi) evaluations of a run -- i.e.
// of all pow calls -- happen inside these trips.
//
// A trip consists of a cheap recurrence (the new psi and step: a dozen operations, needs from the previous trip only ONE BIT,
// "the mass found was larger than the prior mass") and an expensive evaluation (two pow per soil layer).  The mass is a
// decreasing function of psi, so that bit is known without evaluating anything whenever psi is outside a small bracket [a, b]
// around the root.  lg_digit_ff therefore
//   1. finds the root by a safeguarded Newton iteration (the slope comes out of the same two pow values as theta),
//   2. evaluates the mass at a = root - 1.3 w and b = root + 1.3 w (w = tolerance / slope) WITH THE TRIP'S OWN ARITHMETIC and
//      checks that it is above prior + tolerance + margin at a and below prior - tolerance - margin at b,
//   3. replays the walk: every trip runs the recurrence exactly; a trip whose psi lies at or below a (at or above b) is known to
//      find "too wet" ("too dry") and a mass error above the tolerance, so its evaluation is skipped; a trip inside (a, b), or one
//      that meets one of the psi-only exit conditions, is evaluated exactly as the reference does.  The walk ends on an exactly
//      evaluated trip, so theta, psi, the mass error and the trip count are those of the reference bit for bit.
// ~12 evaluations instead of 85, and no heavy tail: a 10^4-trip search costs 10^4 cheap recurrences.
//
// What makes a skipped trip safe:
//   * the computed mass is monotone in psi up to the rounding of two pow and a few products per layer: <= 5e-16 in theta, times
//     the layer thickness (<= 200 cm), three layers: 3e-13 cm.  The bracket test carries a margin of 2e-12 (LG_FF_MARGIN).
//   * the "mass did not change for five trips" exit (count_no_mass_change == 5) looks at |delta_k - delta_{k-1}| < 1e-15.  For two
//     consecutive trips on the same side of the root inside a zone Z where |d mass / d psi| >= g, the change is at least
//     g |psi_k - psi_{k-1}|; if that is >= 1e-12 (LG_FF_NOSTALL) the test is false for certain.  g is a rigorous lower bound:
//     each layer's |d theta / d psi| is unimodal in psi, so its minimum over Z is at an end of Z.  Where that argument does not
//     apply (opposite sides, outside Z, step too small) the counter is carried as an INTERVAL [clo, chi]; if the interval ever
//     straddles 5 while the exit is armed, the search is restarted from its saved state and walked the plain way.
//   * every other exit condition depends on psi, the step and the trip count only and is evaluated exactly on every trip.
// Whenever anything is not certain -- no root in (0, inf), slope 0, bracket test fails, dz < 0, n <= 1 -- the fast path declines
// and the plain walk runs.  tests/test_hostsim.py runs both on every search of every golden case and compares the final records
// bit for bit (-DLGAR_FF_VERIFY); tools/bitexact_sample.py does the same against the unmodified reference over full years.
#ifndef LGAR_SEARCH_CUH
#define LGAR_SEARCH_CUH

#ifdef LGAR_HOSTSIM
#define LG_MFN inline
#define LG_MFN_NOINLINE inline
#else
#define LG_MFN __device__ __forceinline__
#define LG_MFN_NOINLINE __device__ __noinline__
#endif

struct LgSearch {  // one psi search in flight (lgar_theta_mass_balance)
  double psi_loc, factor, psi_prev, delta_mass, delta_mass_prev, new_mass, prior_mass, precip_mass_to_add, theta;
  double dth[LGAR_MAXL + 1], dz[LGAR_MAXL + 1];
  int layer_num, iter, count_no_mass_change;
  bool switched, wanted_to_saturate, aug, aug_extreme, done, no_ff;
  LG_MFN void theta_store(double th) { theta = th; }
};

// one correction search in flight (lgar_theta_mass_balance_correction, lgar.cxx:3262-3435): it works on the
// live front list, so the CORR stage carries the fronts with it
struct LgCorr {
  double psi_loc, factor, psi_prev, delta_mass, delta_mass_prev, new_mass, prior_mass;
  int cur, iter, count_no_mass_change;
  bool switched, aug, aug_extreme, done, no_ff;
  LG_MFN void theta_store(double) {}  // the correction search keeps its thetas in the front list
};

#ifdef LGAR_FF_STATS   // host-side instrumentation (tests/hostsim): how the searches went
struct LgFfStats { long long searches, ff_done, declined, fallback, evals_ff, evals_plain, trips, verify_bad; };
static LgFfStats lg_ff_stats;
#define LG_FF_COUNT(field, n) (lg_ff_stats.field += (n))
#else
#define LG_FF_COUNT(field, n) ((void)0)
#endif

LG_FN void lg_q_note_negative(LgSearch& q) { q.wanted_to_saturate = true; }
LG_FN void lg_q_note_negative(LgCorr&) {}

// The psi / step-size recurrence of one trip (lgar.cxx:2990-3033, :3311-3338).  It needs the outcome of the previous trip only
// as one bit (`up`: the mass found was larger than the prior mass).
template <class Q>
LG_FN void lg_digit_pre(Q& q, bool up) {
  const int first_thresh = LG_MAX_ITER_MBAL / 100, second_thresh = LG_MAX_ITER_MBAL / 10;
  q.iter++;
  if (q.iter > first_thresh && !q.aug) {
    q.factor = q.factor * 100;
    q.aug = true;
  }
  if (q.iter > second_thresh && !q.aug_extreme) {
    q.factor = q.factor * 100;
    q.aug_extreme = true;
  }
  if (up) {
    q.psi_loc += 0.1 * q.factor;
    q.switched = false;
  } else {
    if (!q.switched) {
      q.switched = true;
      q.factor = q.factor * 0.1;
    }
    q.psi_prev = q.psi_loc;
    q.psi_loc -= 0.1 * q.factor;
    if (q.psi_loc < 0.0) lg_q_note_negative(q);
    if (q.psi_loc < 0 && q.psi_prev != 0) q.psi_loc = q.psi_prev * 0.1;
  }
  if (q.psi_loc < 0.0) {
    q.psi_loc = 0.0;
    q.factor = q.factor * 0.5;
  }
}

// the exit conditions that do not look at the mass (lgar.cxx:3063,3078-3086)
template <class Q>
LG_FN bool lg_digit_brk_first(const Q& q) { return fabs(q.psi_loc - q.psi_prev) < 1E-15 && q.factor < 1E-13; }
template <class Q>
LG_FN bool lg_digit_brk_other(const Q& q) {
  const int first_thresh = LG_MAX_ITER_MBAL / 100;
  return (q.iter > LG_MAX_ITER_MBAL) || ((q.psi_loc > LG_PSI_UPPER_LIM) && (q.iter > first_thresh)) || (q.psi_loc <= 0 && q.psi_prev < 0);
}

// the exits of one trip after its evaluation; `gate`: the five-trips-without-change exit is armed
template <class Q>
LG_FN void lg_digit_post(Q& q, bool gate) {
  bool brk = lg_digit_brk_first(q);
  if (!brk) {
    if (fabs(q.delta_mass - q.delta_mass_prev) < 1E-15)
      q.count_no_mass_change++;
    else
      q.count_no_mass_change = 0;
    if (q.count_no_mass_change == 5 && gate) brk = true;
    else brk = lg_digit_brk_other(q);
  }
  if (!brk) q.delta_mass_prev = q.delta_mass;
  if (brk || !(q.delta_mass > LG_MBAL_TOL)) q.done = true;
}

#if !defined(LGAR_FAST_POW)
// theta(h) exactly as lg_theta_from_lnh computes it, and |d theta / d h| from the same two pow values (an estimate: it steers
// the root finder and bounds mass changes from below with a factor 2 in hand, it never enters a result)
LG_FN double lg_theta_and_slope(double h, const LgarLayer& s, double* slope, bool* past_mode) {
  const double t = lg_pow_hot(s.alpha * h, s.n);
  *past_mode = t > s.m;   // |d theta / d h| is unimodal in h with its maximum where (alpha h)^n = m
  const double onep = 1.0 + t;
  const double inv = 1.0 / (lg_pow_hot(onep, s.m));
  double sl = (h > 0.0) ? s.de * s.m * s.n * t * inv / (h * onep) : 0.0;
  if (!(sl >= 0.0) || sl > 1e300) sl = 0.0;
  *slope = sl;
  return inv * s.de + s.theta_r;
}

// Two trips on opposite sides of the root: are their mass errors certainly more than 1e-15 apart?  The root lies in (a, b), so a
// psi at distance d from [a, b] has an error between g d and G (d + b - a) (slope bounds over the zone).
LG_FN bool lg_ff_apart(double x, double y, double a, double b, double g, double G) {
  const double dx = (x <= a) ? a - x : (x >= b ? x - b : 0.0), dy = (y <= a) ? a - y : (y >= b ? y - b : 0.0);
  const double wd = b - a;
  return (g * dx > G * (dy + wd) + 1.0e-12) || (g * dy > G * (dx + wd) + 1.0e-12);
}

#define LG_FF_MARGIN 2.0e-12
#define LG_FF_NOSTALL 1.0e-12
#define LG_FF_MAX_NEWTON 16

// Ev: the mass model of a search.
//   double eval(double psi, double* sl, unsigned* past, double* theta)
//                                         the mass a trip at psi computes (bit for bit; *theta: what the trip stores as theta),
//                                         sl[k] = |d mass/d psi| of term k, bit k of *past: term k is beyond the maximum of its |slope|
//   double term_slope_max(int k)          upper bound of sl[k] over all psi
//   int    terms()                        number of terms of sl
//   bool   monotone()                     the model is one this fast path may be used on
// Returns true when the search is finished (q.done, final record as the plain walk leaves it), false when the fast path
// declined: q is then untouched except q.no_ff.
//
// Shape of the code.  32 searches run side by side in a warp, and an evaluation (two pow per layer, ~600 instructions) is the
// only expensive thing here, so ALL evaluations -- Newton steps, the two bracket points, zone ends, the exactly evaluated trips
// -- go through ONE call site at the top of one loop; between two evaluations each lane runs its own cheap bookkeeping (the
// recurrences of the skipped trips) and comes back to the call site.  Written the straightforward way (evaluate where the need
// arises) the lanes of a warp reach their handful of exact trips at different points of the walk and the warp pays for the union:
// measured on the B200, the six stage kernels fell from 4.3e8 to 1.8e8 column-timesteps/s (profiles/r2_call1_call2_summary.md).
template <class Q, class Ev>
LG_FN bool lg_digit_ff(Q& q, Ev& ev, const bool gate) {
  const Q q0 = q;
  const double prior = q.prior_mass;
  const double TOL = LG_MBAL_TOL;
  q.no_ff = true;
  const double x0 = q.psi_loc;
  if (!ev.monotone() || !(x0 > 0.0) || !(x0 < 1e300)) return false;
  const int nt = ev.terms();
  double sl[Ev::kMaxTerms], slz[Ev::kMaxTerms];
  enum { S_NEWTON, S_BR_A, S_BR_B, S_ZONE_LO, S_ZONE_HI, S_TRIP, S_TRIP_AGAIN };
  enum { W_BEGIN, W_AFTER_EVAL, W_BOOK, W_RECHECK };
  enum { R_RUNNING, R_DONE, R_DECLINED, R_FALLBACK };
  int st = S_NEWTON, wphase = W_BEGIN, result = R_RUNNING;
  double xr = x0;  // where the next evaluation is wanted
  // root finder
  double lo = -1.0, hi = -1.0;  // f(lo) > 0 > f(hi); -1: not known yet
  int newton_it = 0;
  // bracket
  double root = 0.0, w = 0.0, a = 0.0, b = 0.0, fa = 0.0;
  int tries = 0;
  double mul = 1.0;
  // zone Z = [zlo, zhi] with slope bounds g <= |d mass / d psi| <= G on it; only the armed no-change exit needs it.  It is laid
  // round a pair of trips when that pair has to be told apart (2 evaluations), and kept while the walk stays inside it -- after
  // the first turn of the search every later trip lies within one step of the root, so one or two zones serve a whole search.
  double zlo = 0.0, zhi = -1.0, g = 0.0, G = 0.0, nlo = 0.0, nhi = 0.0;
  unsigned pz0 = 0;
  int zones = 0;
  // walk
  int clo = q.count_no_mass_change, chi = q.count_no_mass_change;
  // The record the walk starts from carries a mass the CALLER computed (from stored thetas, possibly clamped), not mass(x0):
  // the first pair of trips is never certified (prev_psi = NaN fails every comparison).
  double prev_psi = lgp_double(0x7ff8000000000000ULL), psi = 0.0;
  int prev_side = (q.new_mass > prior) ? 1 : -1, side = 0, cur_side = 0;
  bool prev_exact = true;  // q.delta_mass_prev is the exact value of the state the walk starts from
  bool skip = false;
  int evals = 0;

  // runs the walk until it needs an evaluation (returns true: st, xr set) or is over (returns false: result set)
  auto advance = [&]() -> bool {
    for (;;) {
      if (wphase == W_BEGIN) {
        lg_digit_pre(q, q.new_mass > prior);
        LG_FF_COUNT(trips, 1);
        psi = q.psi_loc;
        side = (psi <= a) ? 1 : (psi >= b) ? -1 : 0;
        const bool psi_exit = lg_digit_brk_first(q) || lg_digit_brk_other(q);
        skip = side != 0 && !psi_exit;
        cur_side = side;
        if (!skip) {  // a trip inside the bracket, or one that ends the search: evaluated as the reference does
          st = S_TRIP;
          xr = psi;
          wphase = W_AFTER_EVAL;
          return true;
        }
        wphase = W_BOOK;
      }
      if (wphase == W_AFTER_EVAL) {
        cur_side = (q.new_mass > prior) ? 1 : -1;
        wphase = W_BOOK;
      }
      // W_BOOK / W_RECHECK: the no-change counter after this trip, then the end of the trip
      const bool first = !skip && lg_digit_brk_first(q);
      bool brk = first;
      if (gate && !first) {
        if (!skip && prev_exact) {
          if (fabs(q.delta_mass - q.delta_mass_prev) < 1E-15) {
            clo++;
            chi++;
          } else {
            clo = chi = 0;
          }
        } else {
          // at least one of the two mass errors is not known: can they be within 1e-15 of each other?  Not if both trips lie in a
          // zone with slope bounds and are far enough apart (same side), or at clearly different distances from the root
          // (opposite sides).  Both errors are above the tolerance here (a skipped trip by the bracket, an exact one because the
          // search goes on), so on the same side |delta_k - delta_{k-1}| is the change of the mass itself.
          bool apart = false;
          if (prev_psi == prev_psi && (skip || q.delta_mass > TOL)) {
            const bool inz = psi >= zlo && psi <= zhi && prev_psi >= zlo && prev_psi <= zhi;
            if (inz) apart = (cur_side == prev_side) ? (g * fabs(psi - prev_psi) >= LG_FF_NOSTALL) : lg_ff_apart(psi, prev_psi, a, b, g, G);
            if (!apart && wphase == W_BOOK && zones < 8 && !(inz && !(G > 1.5 * g))) {
              // lay a new zone round this pair and the bracket, with room for one more step in the direction of the walk
              const double step = fabs(psi - prev_psi);
              const double pad_lo = (psi < prev_psi ? 1.01 : 0.11) * step, pad_hi = (psi > prev_psi ? 1.01 : 0.11) * step;
              nlo = fmin(fmin(psi, prev_psi), a) - pad_lo;
              nhi = fmax(fmax(psi, prev_psi), b) + pad_hi;
              if (!(nlo > 0.5 * fmin(fmin(psi, prev_psi), a))) nlo = 0.5 * fmin(fmin(psi, prev_psi), a);
              if (nlo > 0.0 && nhi < 1e300) {
                st = S_ZONE_LO;
                xr = nlo;
                wphase = W_RECHECK;  // where the walk resumes once both ends of the zone are evaluated
                return true;
              }
            }
          }
          if (apart) {
            clo = chi = 0;
          } else {
            clo = 0;
            chi++;
          }
        }
        if (!skip && clo == 5 && chi == 5) {
          brk = true;
        } else if (clo <= 5 && chi >= 5) {  // the no-change exit can neither be ruled out nor confirmed: walk the plain way
          result = R_FALLBACK;
          return false;
        }
      }
      wphase = W_BEGIN;
      if (skip) {
        q.new_mass = (side > 0) ? prior + 1.0 : prior - 1.0;  // only its order relative to the prior mass is read
        prev_psi = psi;
        prev_side = side;
        prev_exact = false;
        continue;
      }
      if (!brk) brk = !first && lg_digit_brk_other(q);
      if (!brk) q.delta_mass_prev = q.delta_mass;
      if (brk || !(q.delta_mass > TOL)) {
        q.done = true;
        q.count_no_mass_change = clo;
        result = R_DONE;
        return false;
      }
      prev_psi = psi;
      prev_side = cur_side;
      prev_exact = true;
    }
  };

  for (;;) {
    // ---- the one evaluation site ----
    unsigned pm = 0;
    double th = 0.0;
    const double m = ev.eval(xr, st == S_ZONE_HI ? slz : sl, &pm, &th);
    evals++;
    bool more = true;
    if (st == S_NEWTON) {
      // root of f(psi) = mass(psi) - prior by a safeguarded Newton iteration, started at the current psi
      const double fx = m - prior;
      double sx = 0.0;
      for (int k = 0; k < nt; k++) sx += sl[k];
      if (!(fx == fx)) {
        result = R_DECLINED;
        more = false;
      } else if (fabs(fx) <= 0.25 * TOL) {
        root = xr;
        w = (sx > 0.0) ? TOL / sx : 0.0;
        a = root - 1.3 * w;
        b = root + 1.3 * w;
        if (!(sx > 0.0) || !(a > 0.0) || !(a < root) || !(b > root)) {
          result = R_DECLINED;
          more = false;
        } else {
          st = S_BR_A;
          xr = a;
        }
      } else {
        if (fx > 0.0) lo = xr; else hi = xr;
        double xn = (sx > 0.0) ? xr + fx / sx : (fx > 0.0 ? 2.0 * xr : 0.5 * xr);  // f decreases: f' = -sx
        if (xn > 4.0 * xr) xn = 4.0 * xr;
        if (xn < 0.25 * xr) xn = 0.25 * xr;
        if (lo > 0.0 && hi > 0.0 && !(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        newton_it++;
        if (!(xn == xn) || newton_it >= LG_FF_MAX_NEWTON || xn < 1e-300 || xn > 1e300 || (lo > 0.0 && hi > 0.0 && !(xn > lo && xn < hi))) {
          result = R_DECLINED;  // no root in reach (the search is heading for psi = 0 or for the upper limit), or a bracket closed
          more = false;         // to neighbouring doubles without meeting a quarter of the tolerance
        } else {
          xr = xn;
        }
      }
    } else if (st == S_BR_A) {
      // the bracket: mass error above the tolerance (with margin) at a, below minus the tolerance at b
      fa = m - prior;
      st = S_BR_B;
      xr = b;
    } else if (st == S_BR_B) {
      const double fb = m - prior;
      if ((fa >= TOL + LG_FF_MARGIN) && (fb <= -(TOL + LG_FF_MARGIN))) {
        more = advance();
      } else {
        tries++;
        mul *= 2.0;
        a = root - 1.3 * mul * w;
        b = root + 1.3 * mul * w;
        if (tries >= 3 || !(a > 0.0)) {
          result = R_DECLINED;
          more = false;
        } else {
          st = S_BR_A;
          xr = a;
        }
      }
    } else if (st == S_ZONE_LO) {
      pz0 = pm;
      st = S_ZONE_HI;
      xr = nhi;
    } else if (st == S_ZONE_HI) {
      g = 0.0;
      G = 0.0;
      for (int k = 0; k < nt; k++) {
        g += fmin(sl[k], slz[k]);
        G += (((pz0 ^ pm) >> k) & 1u) ? ev.term_slope_max(k) : fmax(sl[k], slz[k]);
      }
      g *= 0.9;
      G *= 1.1;
      zlo = nlo;
      zhi = nhi;
      zones++;
      if (!skip && Ev::kEvalHasSideEffects) {
        // the model's evaluation writes into the live state (the correction search sets theta and psi of its chain of fronts): the
        // zone ends have just overwritten what the exactly evaluated trip left there -- evaluate the trip once more
        st = S_TRIP_AGAIN;
        xr = psi;
      } else {
        more = advance();
      }
    } else if (st == S_TRIP_AGAIN) {
      more = advance();
    } else {  // S_TRIP: the evaluation of the trip at q.psi_loc
      q.theta_store(th);
      q.new_mass = m;
      q.delta_mass = fabs(m - prior);
      more = advance();
    }
    if (!more) break;
  }
  LG_FF_COUNT(evals_ff, evals);
  if (result == R_DONE) {
    LG_FF_COUNT(ff_done, 1);
    return true;
  }
  if (result == R_DECLINED) LG_FF_COUNT(declined, 1); else LG_FF_COUNT(fallback, 1);
  q = q0;
  q.no_ff = true;
  return false;
}
#endif  // !LGAR_FAST_POW

// ---- lgar_theta_mass_balance: the mass of the layers above and of the front's own layer at one psi ----------------------
LG_FN void lg_search_eval(LgSearch& q, const LgarColumnClass& cls) {
  const int layer_num = q.layer_num;
  const double lnpsi = lg_lnh(q.psi_loc);
  q.theta = lg_theta_from_lnh(lnpsi, cls.lay[layer_num]);
  double mass_layers = 0.0;
  mass_layers += q.dz[layer_num] * (q.theta - q.dth[layer_num]);
  for (int k = 1; k < layer_num; k++) {
    double theta_layer = lg_theta_from_lnh(lnpsi, cls.lay[k]);
    mass_layers += q.dz[k] * (theta_layer - q.dth[k]);
  }
  q.new_mass = mass_layers;
  q.delta_mass = fabs(q.new_mass - q.prior_mass);
}
// the initial record of a search (lgar.cxx:2961-2983)
LG_FN void lg_search_setup(LgSearch& q, int layer_num, double psi_cm, double new_mass, double prior_mass, double precip_mass_to_add) {
  q.layer_num = layer_num;
  q.psi_loc = psi_cm;
  q.new_mass = new_mass;
  q.prior_mass = prior_mass;
  q.precip_mass_to_add = precip_mass_to_add;
  q.delta_mass = fabs(new_mass - prior_mass);
  q.factor = fmax(1.0, psi_cm / 10.0);
  if (psi_cm > 1.E4) q.factor = psi_cm;
  q.switched = false;
  q.theta = 0;
  q.psi_prev = psi_cm;
  q.delta_mass_prev = q.delta_mass;
  q.count_no_mass_change = 0;
  q.wanted_to_saturate = false;
  q.iter = 0;
  q.aug = false;
  q.aug_extreme = false;
  q.done = false;
  q.no_ff = false;
}
LG_FN void lg_search_pre(LgSearch& q, bool up) { lg_digit_pre(q, up); }
LG_FN void lg_search_post(LgSearch& q) { lg_digit_post(q, q.precip_mass_to_add < 1.E-12); }
// one trip; sets q.done when the loop would exit
LG_FN void lg_search_iter(LgSearch& q, const LgarColumnClass& cls) {
  lg_digit_pre(q, q.new_mass > q.prior_mass);
  lg_search_eval(q, cls);
  lg_search_post(q);
  LG_FF_COUNT(evals_plain, 1);
}

#if !defined(LGAR_FAST_POW)
struct LgSearchModel {
  static constexpr int kMaxTerms = LGAR_MAXL + 1;
  static constexpr bool kEvalHasSideEffects = false;
  const LgSearch* q;
  const LgarColumnClass* cls;
  LG_MFN bool monotone() const {
    for (int k = 1; k <= q->layer_num; k++)
      if (!(q->dz[k] >= 0.0) || !(cls->lay[k].n > 1.0) || !(cls->lay[k].alpha > 0.0) || !(cls->lay[k].de > 0.0)) return false;
    return true;
  }
  LG_MFN int terms() const { return q->layer_num; }
  LG_MFN double term_slope_max(int k) const { const int l = k == 0 ? q->layer_num : k; return q->dz[l] * cls->lay[l].slope_max; }
  LG_MFN_NOINLINE double eval(double psi, double* sl, unsigned* past, double* theta_out) const {   // lg_search_eval at psi
    const int layer_num = q->layer_num;
    double s;
    bool pm;
    unsigned pb = 0;
    const double theta = lg_theta_and_slope(psi, cls->lay[layer_num], &s, &pm);
    *theta_out = theta;
    pb |= pm ? 1u : 0u;
    sl[0] = q->dz[layer_num] * s;
    double mass_layers = 0.0;
    mass_layers += q->dz[layer_num] * (theta - q->dth[layer_num]);
    for (int k = 1; k < layer_num; k++) {
      const double theta_layer = lg_theta_and_slope(psi, cls->lay[k], &s, &pm);
      pb |= pm ? (1u << k) : 0u;
      sl[k] = q->dz[k] * s;
      mass_layers += q->dz[k] * (theta_layer - q->dth[k]);
    }
    *past = pb;
    return mass_layers;
  }
};
#endif

// A search in flight: through the fast path if `use_ff` (from whatever trip the walk has reached), then at most `quantum` plain
// trips of whatever is left.  Where the fast path pays: it does ~10 evaluations where the plain walk does ~70, but the lanes of a
// warp walking plainly are CONVERGED and compacted (the SEARCH stage kernel runs at ~65 % of the FP64 peak), while the fast
// path's bookkeeping diverges -- measured on the B200, bulk rounds are 2.1 x SLOWER with it (profiles/r2_call1_call2_summary.md).
// It is a latency tool: the engine turns it on when few slots are in flight (drop-out rounds, finishing kernel), where a round
// lasts as long as its longest search and a 10^4-trip monster is 10^4 cheap recurrences instead of 10^4 evaluations.
LG_FN void lg_search_some(LgSearch& q, const LgarColumnClass& cls, int quantum, bool use_ff) {
#if !defined(LGAR_FAST_POW) && !defined(LGAR_NO_FF)
  if (use_ff && !q.no_ff) {  // from whatever trip the walk has reached: the record is a complete state of it
    LgSearchModel ev{&q, &cls};
    if (lg_digit_ff(q, ev, q.precip_mass_to_add < 1.E-12)) return;
  }
#endif
  for (int it = 0; it < quantum && !q.done; it++) lg_search_iter(q, cls);
}

// the whole search to q.done
LG_FN_NOINLINE void lg_search_run(LgSearch& q, const LgarColumnClass& cls, bool use_ff) {
  LG_FF_COUNT(searches, 1);
#ifdef LGAR_FF_VERIFY
  LgSearch ref = q;
  while (!ref.done) lg_search_iter(ref, cls);
  const bool fresh = use_ff && !q.no_ff;
#endif
  while (!q.done) lg_search_some(q, cls, 1 << 30, use_ff);
#ifdef LGAR_FF_VERIFY
  if (fresh && !(lgp_bits(ref.psi_loc) == lgp_bits(q.psi_loc) && lgp_bits(ref.theta) == lgp_bits(q.theta) && lgp_bits(ref.delta_mass) == lgp_bits(q.delta_mass) &&
                 lgp_bits(ref.new_mass) == lgp_bits(q.new_mass) && ref.iter == q.iter && ref.wanted_to_saturate == q.wanted_to_saturate))
    LG_FF_COUNT(verify_bad, 1);
#endif
}

// ---- lgar_theta_mass_balance_correction (lgar.cxx:3262-3435) --------------------------------------------------------------
// The evaluation of one trip on the live front list: the fronts that take the new psi are `cur`, the to-bottom fronts chained
// below it and the first free front after such a chain (lgar.cxx:3340-3375; every psi in the chain is a copy of psi_loc), then
// the mass of the whole list (lgar_calc_mass_bal, lgar.cxx:2632-2668).
LG_FN int lg_corr_chain_end(const LgFronts& f, int cur) {
  int last = cur;
  int nx = cur + 1;
  while (nx < f.n && f.tb[nx]) {
    last = nx;
    nx++;
  }
  if (nx < f.n && f.tb[last]) last = nx;  // tb[nx] is false here
  return last;
}
LG_FN double lg_corr_list_mass(const LgFronts& f, const LgarColumnClass& cls) {
  const double* cum = cls.cum;
  double sum = 0.0;
#pragma unroll 1
  for (int i = 0; i < f.n; i++) {
    double base = cum[f.layer[i] - 1];
    if (i + 1 < f.n) {
      if (f.layer[i + 1] == f.layer[i])
        sum += (f.depth[i] - base) * (f.theta[i] - f.theta[i + 1]);
      else
        sum += (f.depth[i] - base) * f.theta[i];
    } else {
      sum += f.theta[i] * (f.depth[i] - base);
    }
  }
  return sum;
}
LG_FN void lg_corr_eval(LgCorr& k, LgFronts& f, const LgarColumnClass& cls) {
  const int cur = k.cur;
  const double lnpsi = lg_lnh(k.psi_loc);
  const int last = lg_corr_chain_end(f, cur);
#pragma unroll 1
  for (int j = cur; j <= last; j++) {
    f.psi[j] = k.psi_loc;
    f.theta[j] = lg_theta_from_lnh(lnpsi, cls.lay[f.layer[j]]);
  }
  k.new_mass = lg_corr_list_mass(f, cls);
  k.delta_mass = fabs(k.new_mass - k.prior_mass);
}
// one trip of the loop
LG_FN void lg_corr_iter(LgCorr& k, LgFronts& f, const LgarColumnClass& cls) {
  lg_digit_pre(k, k.new_mass > k.prior_mass);
  lg_corr_eval(k, f, cls);
  lg_digit_post(k, true);  // the five-trips-without-change exit is always armed here (lgar.cxx:3418)
}

#if !defined(LGAR_FAST_POW)
// The mass model of the correction search for the fast path of lgar_search.cuh.  The list mass is linear in the thetas of the
// chain fronts; front j of the chain carries the coefficient (depth_j - top of its layer), minus the same of front j-1 when that
// one is in the same layer.  Terms are grouped by soil layer (one unimodal |d theta / d psi| per layer).
struct LgCorrModel {
  static constexpr int kMaxTerms = LGAR_MAXL + 1;
  static constexpr bool kEvalHasSideEffects = true;   // eval() writes theta and psi of the chain into the front list
  LgCorr* k;
  LgFronts* f;
  const LgarColumnClass* cls;
  int last;
  double coef[LGAR_MAXL + 1];
  bool ok;
  LG_MFN void init() {
    const LgFronts& F = *f;
    last = lg_corr_chain_end(F, k->cur);
    ok = true;
    for (int l = 0; l <= LGAR_MAXL; l++) coef[l] = 0.0;
    for (int j = k->cur; j <= last; j++) {
      const int l = F.layer[j];
      const LgarLayer& s = cls->lay[l];
      double c = F.depth[j] - cls->cum[l - 1];
      if (j > 0 && F.layer[j - 1] == l) c -= F.depth[j - 1] - cls->cum[l - 1];
      if (!(c >= 0.0) || !(s.n > 1.0) || !(s.alpha > 0.0) || !(s.de > 0.0)) ok = false;
      coef[l] += c;
    }
  }
  LG_MFN bool monotone() const { return ok; }
  LG_MFN int terms() const { return LGAR_MAXL + 1; }
  LG_MFN double term_slope_max(int t) const { return coef[t] * cls->lay[t].slope_max; }
  LG_MFN_NOINLINE double eval(double psi, double* sl, unsigned* past, double* theta_out) const {   // lg_corr_eval at psi
    LgFronts& F = *f;
    *theta_out = 0.0;
    for (int l = 0; l <= LGAR_MAXL; l++) sl[l] = 0.0;
    unsigned pb = 0;
#pragma unroll 1
    for (int j = k->cur; j <= last; j++) {
      const int l = F.layer[j];
      double s;
      bool pm;
      F.psi[j] = psi;
      F.theta[j] = lg_theta_and_slope(psi, cls->lay[l], &s, &pm);
      sl[l] = coef[l] * s;
      pb |= pm ? (1u << l) : 0u;
    }
    *past = pb;
    return lg_corr_list_mass(F, *cls);
  }
};
#endif

// A freshly initialised correction search (lg_corr_init returned true) through the fast path, then at most `quantum` plain
// trips of whatever is left.
LG_FN void lg_corr_some(LgCorr& k, LgFronts& f, const LgarColumnClass& cls, int quantum, bool use_ff) {
#if !defined(LGAR_FAST_POW) && !defined(LGAR_NO_FF)
  if (use_ff && !k.no_ff) {
    LgCorrModel ev;
    ev.k = &k;
    ev.f = &f;
    ev.cls = &cls;
    ev.init();
    if (lg_digit_ff(k, ev, true)) return;
  }
#endif
  for (int it = 0; it < quantum && !k.done; it++) lg_corr_iter(k, f, cls);
}
// the whole correction search to k.done
LG_FN_NOINLINE void lg_corr_run(LgCorr& k, LgFronts& f, const LgarColumnClass& cls, bool use_ff) {
#ifdef LGAR_FF_VERIFY
  LgCorr ref = k;
  static LgFronts fref;
  fref = f;
  while (!ref.done) lg_corr_iter(ref, fref, cls);
  const bool fresh = use_ff && !k.no_ff;
  LG_FF_COUNT(searches, 1);
#endif
  while (!k.done) lg_corr_some(k, f, cls, 1 << 30, use_ff);
#ifdef LGAR_FF_VERIFY
  bool same = lgp_bits(ref.psi_loc) == lgp_bits(k.psi_loc) && lgp_bits(ref.delta_mass) == lgp_bits(k.delta_mass) && lgp_bits(ref.new_mass) == lgp_bits(k.new_mass) && ref.iter == k.iter;
  for (int i = 0; i < f.n; i++) same = same && lgp_bits(fref.theta[i]) == lgp_bits(f.theta[i]) && lgp_bits(fref.psi[i]) == lgp_bits(f.psi[i]);
  if (fresh && !same) LG_FF_COUNT(verify_bad, 1);
#endif
}

#endif
