/*
 * tsc_sim_ref.c — CPU ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Scalar, sequential restatement of one control step of the traffic-signal environment of
 * cts198859/deeprl_signal_control:
 *     TrafficSimulator.step()            reference envs/env.py:566-631
 *       _set_phase / _get_node_phase     reference envs/env.py:455-459, 128-152
 *       _simulate -> traci.simulationStep  reference envs/env.py:461-471   (SUMO, NOT in the repo)
 *       _measure_state_step              reference envs/env.py:369-407, 439-442
 *       _measure_reward_step             reference envs/env.py:325-367
 *       _get_state                       reference envs/env.py:163-205
 *       reward shaping                   reference envs/env.py:591-631
 *
 * PARITY STATUS
 *   - control protocol, observation, reward, shaping: pinned against the reference's own
 *     Python (envs/env.py executed with a fake TraCI connection, tests/golden/gen_env_golden.py).
 *   - vehicle dynamics ("simulationStep"): the arithmetic lives in Eclipse SUMO >= 1.1.0
 *     (README.md:26), which is absent from /root/reference and from this image.  What follows
 *     restates SUMO's published Krauss model (MSCFModel_Krauss / MSCFModel: Euler update,
 *     maximumSafeStopSpeedEuler, brakeGap, freeSpeed, dawdle2) plus the documented
 *     simplifications of DESIGN.md §3.  => "parity unpinned" against SUMO; bit-exact parity is
 *     claimed only between this file and the CUDA kernel (deeprl_signal_control_b200/csrc).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.
 *
 * All floating-point work is IEEE binary32 with no FMA contraction (-ffp-contract=off here,
 * -fmad=false in nvcc) and only + - * / sqrt floor, so results are bit-identical to the GPU.
 */
#include <math.h>
#include <pthread.h>
#include <alloca.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tsc.h"

#define INF_SPEED 1.0e9f
#define F_CROSS 1
#define F_ARRIVE 2

typedef struct {
  uint32_t xv; /* position : 16 (1/64 m) | speed : 16 (1/1024 m/s), see tsc.h "vehicle record" */
  uint32_t m0; /* wait:10 | hop:6 | route:8 | sfq:8 */
} veh_t;   /* 8 bytes; trip statistics (depart, total wait) are not carried on the hot path */

/* fixed-point state: both scales are powers of two, so unpacking is exact */
#define XQ_SCALE 64.0f
#define XQ_INV 0.015625f
#define VQ_SCALE 1024.0f
#define VQ_INV 0.0009765625f
static inline float veh_x(uint32_t xv) { return (float)(xv & 0xffffu) * XQ_INV; }
static inline float veh_v(uint32_t xv) { return (float)(xv >> 16) * VQ_INV; }
static inline uint32_t pack_xv(float x, float v) {
  /* positions are truncated (a vehicle that stops 1 mm short of a stop line must stay short of it), speeds rounded */
  int xq = (int)(x * XQ_SCALE), vq = (int)(v * VQ_SCALE + 0.5f);
  if (xq < 0) xq = 0;
  if (xq > 65535) xq = 65535;
  if (vq < 0) vq = 0;
  if (vq > 65535) vq = 65535;
  return (uint32_t)xq | ((uint32_t)vq << 16);
}

typedef struct {
  veh_t* ring;          /* [n_slots] */
  int32_t* head;        /* [n_lanes] */
  int32_t* cnt;         /* [n_lanes] */
  int32_t* prev_action; /* [n_nodes] */
  int32_t* cur_action;  /* [n_nodes] */
  int32_t* backlog;     /* [n_src]   */
  int32_t cur_sec;
  uint32_t seed_lo, seed_hi;
  int32_t n_departed, n_arrived;
  /* measurement of the last step */
  int32_t* det_veh;     /* [n_det] */
  int32_t* det_halt;
  int32_t* det_wait;
  /* record mode (evaluation runs, envs/env.py:498-542): trip word per ring slot, parallel to `ring`
   * (depart:12 | total wait s:12 | wait episodes:8), and the arrival log = tripinfo rows
   * {depart:12 | arrival:12 | route:8, wait s:16 | wait episodes:16}; NULL when record mode is off */
  uint32_t* trip;       /* [n_slots] */
  uint32_t* trip_log;   /* [trip_cap][2] */
  int32_t trip_cnt, trip_cap;
} replica_t;

typedef struct ref_sim {
  tsc_net net;
  tsc_cfg cfg;
  int32_t R;
  int32_t train_mode;
  replica_t* rep;
} ref_sim;

/* ------------------------------------------------------------------------------------------ */
/* counter-based random numbers: identical integer hash on CPU and GPU                         */
static inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
  return h;
}
/* key of one simulated second of one replica (two rounds), then ONE round per draw (entity b, index c) */
static inline uint32_t rng_key(uint32_t s0, uint32_t s1, uint32_t a) {
  return mix32(mix32(s0 ^ (a * 0x9E3779B1U)) ^ s1);
}
static inline uint32_t rng_draw(uint32_t key, uint32_t b, uint32_t c) {
  return mix32(key ^ (b * 0x85EBCA77U + c * 0xC2B2AE3DU));
}
static inline uint32_t rng_u32(uint32_t s0, uint32_t s1, uint32_t a, uint32_t b, uint32_t c) {
  return rng_draw(rng_key(s0, s1, a), b, c);
}
static inline float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

/* ------------------------------------------------------------------------------------------ */
/* Krauss car-following, Euler update, dt = 1 s (SUMO MSCFModel.cpp, restated)                 */
/* ib = 1 / b is formed once (float division) and multiplied: one IEEE division less per call, same on CPU and GPU */
static inline float brake_gap(float v, float b, float ib) {
  int steps = (int)(v * ib);
  float fs = (float)steps;
  float t1 = fs * v;
  float t2 = b * fs;
  float t3 = fs + 1.0f;
  float t4 = t2 * t3;
  float t5 = t4 * 0.5f;
  return t1 - t5;
}
static inline float stop_speed(float gap, float b, float ib, float tau) {
  float g = gap - 0.001f;
  if (g < 0.0f) return 0.0f;
  float q = (2.0f * g) * ib;
  q = q - tau;
  float tt = tau * tau;
  float disc = 1.0f + 4.0f * (q + tt);
  float sq = sqrtf(disc);
  float n = floorf(0.5f - (tau + sq * -0.5f));
  float h1 = 0.5f * n;
  h1 = h1 * (n - 1.0f);
  h1 = h1 * b;
  float h2 = n * b;
  h2 = h2 * tau;
  float h = h1 + h2;
  float r = (g - h) / (n + tau);
  return n * b + r;
}
static inline float follow_speed(float gap, float v_lead, float b, float ib, float tau) {
  return stop_speed(gap + brake_gap(v_lead, b, ib), b, ib, tau);
}
static inline float free_speed(float dist, float target, float b) {
  if (dist < target) return target;
  float bb = b + 2.0f * target;
  float disc = bb * bb + (8.0f * b) * dist;
  float y = ((sqrtf(disc) - b) * 0.5f - target) / b;
  if (y < 0.0f) y = 0.0f;
  float yf = floorf(y);
  float eg = (yf * yf + yf) * 0.5f;
  eg = eg * b;
  eg = eg + yf * target;
  if (y > yf) eg = eg + target;
  float rem = dist - eg;
  if (rem < 0.0f) rem = 0.0f;
  float res = rem / (yf + 1.0f);
  res = res + yf * b;
  return res + target;
}

/* ------------------------------------------------------------------------------------------ */
static inline veh_t* veh_at(const tsc_net* n, replica_t* r, int lane, int rank) {
  int cap = n->lane_cap[lane];
  int idx = r->head[lane] + rank;
  if (idx >= cap) idx -= cap;
  return &r->ring[n->lane_slot0[lane] + idx];
}
#define M0_WAIT(m) ((m) & 1023u)
#define M0_HOP(m) (((m) >> 10) & 63u)
#define M0_ROUTE(m) (((m) >> 16) & 255u)
#define M0_SFQ(m) ((m) >> 24)
#define T1_DEPART(t) ((t) & 4095u)
#define T1_WAIT(t) (((t) >> 12) & 4095u)
#define T1_WCNT(t) ((t) >> 24)

/* signal state of every node for one sub-step: reference envs/env.py:128-152 */
static void node_signal(const ref_sim* s, const replica_t* r, int yellow_phase, uint32_t* open,
                        uint32_t* major, uint32_t* ymask) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < n->n_nodes; ++i) {
    int a = r->cur_action[i], p = r->prev_action[i];
    uint32_t g1 = n->node_green[i * n->max_phases + a];
    uint32_t m1 = n->node_major[i * n->max_phases + a];
    open[i] = g1; major[i] = m1; ymask[i] = 0;
    if (yellow_phase && p >= 0 && p != a) {
      uint32_t g0 = n->node_green[i * n->max_phases + p];
      uint32_t sw_red = g0 & ~g1;     /* 'G'/'g' -> 'r' : shown yellow          */
      uint32_t sw_green = ~g0 & g1;   /* 'r' -> 'G'/'g' : held red during yellow */
      if (sw_red) {
        ymask[i] = sw_red;
        open[i] = g1 & ~sw_green;
        major[i] = m1 & ~sw_green;
      }
    }
  }
}

/* one simulated second of one replica == traci.simulationStep(), envs/env.py:464 */
static void substep(ref_sim* s, replica_t* r, int yellow_phase, float* vnew, float* xnew,
                    uint8_t* flag, uint8_t* hblk, float* head_lim, uint32_t* approach, uint32_t* open,
                    uint32_t* major, uint32_t* ymask, uint8_t* accepted, int32_t* cnt_add) {
  const tsc_net* n = &s->net;
  const tsc_cfg* c = &s->cfg;
  const int L = n->n_lanes;
  const uint32_t t_abs = (uint32_t)r->cur_sec;
  const uint32_t key = rng_key(r->seed_lo, r->seed_hi, t_abs);
  const float ib = 1.0f / c->decel;
  node_signal(s, r, yellow_phase, open, major, ymask);

  /* A1: which links have an approaching head vehicle with a green light */
  for (int i = 0; i < n->n_nodes; ++i) approach[i] = 0;
  for (int l = 0; l < L; ++l) {
    if (r->cnt[l] == 0) continue;
    veh_t* h = veh_at(n, r, l, 0);
    int link = n->route_link[M0_ROUTE(h->m0) * n->max_hops + M0_HOP(h->m0)];
    if (link < 0) continue;
    int node = n->link_node[link];
    if (node < 0) continue;
    uint32_t bit = 1u << n->link_tlidx[link];
    float d = n->lane_len[l] - veh_x(h->xv);
    if ((open[node] & bit) && d <= 3.0f * veh_v(h->xv) + 7.5f) approach[node] |= bit;
  }
  /* A2: speed limit of each lane's head vehicle from the junction ahead */
  for (int l = 0; l < L; ++l) {
    head_lim[l] = INF_SPEED;
    hblk[l] = 0;
    if (r->cnt[l] == 0) continue;
    veh_t* h = veh_at(n, r, l, 0);
    uint32_t route = M0_ROUTE(h->m0), hop = M0_HOP(h->m0);
    int link = n->route_link[route * n->max_hops + hop];
    if (link < 0) continue; /* arrival lane */
    float d = n->lane_len[l] - veh_x(h->xv);
    int node = n->link_node[link];
    int blocked = 0;
    if (node >= 0 && (int)M0_WAIT(h->m0) < c->teleport_sec) {
      uint32_t bit = 1u << n->link_tlidx[link];
      if (ymask[node] & bit) {
        blocked = brake_gap(veh_v(h->xv), c->decel, ib) <= d;
      } else if (!(open[node] & bit)) {
        blocked = 1;
      } else {
        uint32_t foes = n->link_merge[link];
        if (!(major[node] & bit)) foes |= n->link_cross[link];
        if (approach[node] & foes) blocked = 1;
      }
    }
    if (blocked) { head_lim[l] = stop_speed(d, c->decel, ib, c->tau); hblk[l] = 1; continue; }
    float lim = INF_SPEED;
    float lv = n->link_vmax[link];
    if (lv < 1.0e8f) lim = free_speed(d, lv, c->decel);
    int nl = n->route_lane[route * n->max_hops + hop + 1];
    if (r->cnt[nl] > 0) {
      veh_t* t = veh_at(n, r, nl, r->cnt[nl] - 1);
      float gap = d + (veh_x(t->xv) - c->veh_len);
      gap = gap - c->min_gap;
      float fs = follow_speed(gap, veh_v(t->xv), c->decel, ib, c->tau);
      if (fs < lim) lim = fs;
    }
    head_lim[l] = lim;
  }
  /* B: every vehicle plans and executes its move from the OLD state */
  for (int l = 0; l < L; ++l) {
    int base = n->lane_slot0[l];
    float Ll = n->lane_len[l];
    for (int k = 0; k < r->cnt[l]; ++k) {
      veh_t* v = veh_at(n, r, l, k);
      int slot = (int)(v - r->ring);
      (void)base;
      const float vx = veh_x(v->xv), vv = veh_v(v->xv);
      float sf = 0.5f + (float)M0_SFQ(v->m0) * (1.0f / 256.0f);
      float vmax = n->lane_vmax[l] * sf;
      float vfree = vv + c->accel;
      if (vmax < vfree) vfree = vmax;
      float vsafe;
      if (k == 0) {
        vsafe = head_lim[l];
      } else {
        veh_t* ld = veh_at(n, r, l, k - 1);
        float gap = veh_x(ld->xv) - c->veh_len;
        gap = gap - vx;
        gap = gap - c->min_gap;
        vsafe = follow_speed(gap, veh_v(ld->xv), c->decel, ib, c->tau);
      }
      float vnm = vfree < vsafe ? vfree : vsafe;
      float vmin = vv - c->decel;
      if (vmin < 0.0f) vmin = 0.0f;
      if (vnm < vmin) vmin = vnm;
      float u = u01(rng_draw(key, (uint32_t)l, (uint32_t)k));
      float basev = vnm < c->accel ? vnm : c->accel;
      float vd = vnm - (c->sigma * basev) * u;
      float vn = vd > vmin ? vd : vmin;
      float xn = vx + vn;
      uint8_t f = 0;
      if (xn >= Ll) {
        if (k == 0 && !hblk[l]) {
          int link = n->route_link[M0_ROUTE(v->m0) * n->max_hops + M0_HOP(v->m0)];
          f = link < 0 ? F_ARRIVE : F_CROSS;
        } else { /* a lane discharges at most one vehicle per second; a head vehicle whose stop line is closed never
                  * passes it (tau < 1 s makes the Euler stop speed overshoot: SUMO's "emergency stop at the end of
                  * the lane") */
          xn = Ll - 0.01f;
          vn = xn - vx;
          if (vn < 0.0f) { vn = 0.0f; xn = vx; }
        }
      }
      vnew[slot] = vn; xnew[slot] = xn; flag[slot] = f;
    }
  }
  /* commit B (sequential code needs the old state untouched until here) */
  for (int l = 0; l < L; ++l) {
    for (int k = 0; k < r->cnt[l]; ++k) {
      veh_t* v = veh_at(n, r, l, k);
      int slot = (int)(v - r->ring);
      const float vn = vnew[slot];      /* waiting is decided on the computed speed, the record stores it rounded */
      v->xv = pack_xv(xnew[slot], vn);
      uint32_t w = M0_WAIT(v->m0);
      if (r->trip && vn < 0.1f) { /* tripinfo waitingTime / waitingCount */
        uint32_t t1 = r->trip[slot];
        uint32_t wt = T1_WAIT(t1), wc = T1_WCNT(t1);
        if (wt < 4095u) wt++;
        if (w == 0 && wc < 255u) wc++;
        r->trip[slot] = T1_DEPART(t1) | (wt << 12) | (wc << 24);
      }
      if (vn < 0.1f) {
        if (w < 1023u) w++;
      } else {
        w = 0;
      }
      v->m0 = (v->m0 & ~1023u) | w;
    }
  }
  /* C: junction transfers, one destination lane at a time, sources in merge-priority order */
  for (int l = 0; l < L; ++l) { accepted[l] = 0; cnt_add[l] = 0; }
  for (int t = 0; t < L; ++t) {
    int cur = r->cnt[t];
    int have_tail = cur > 0;
    float tail_x = have_tail ? veh_x(veh_at(n, r, t, cur - 1)->xv) : 0.0f;
    for (int q = n->lane_inl_off[t]; q < n->lane_inl_off[t + 1]; ++q) {
      int link = n->lane_inl[q];
      int src = n->link_from[link];
      if (r->cnt[src] == 0) continue;
      veh_t* h = veh_at(n, r, src, 0);
      int hslot = (int)(h - r->ring);
      if (flag[hslot] != F_CROSS) continue;
      uint32_t route = M0_ROUTE(h->m0), hop = M0_HOP(h->m0);
      if (n->route_link[route * n->max_hops + hop] != link) continue;
      if (n->route_lane[route * n->max_hops + hop + 1] != t) continue;
      if (cur >= n->lane_cap[t]) continue; /* refused: ring full */
      float x = veh_x(h->xv) - n->lane_len[src];
      if (have_tail) {
        float lim = tail_x - c->veh_len;
        lim = lim - c->min_gap;
        if (x > lim) x = lim;
      }
      if (x < 0.0f) continue; /* refused: no room behind the tail */
      int cap = n->lane_cap[t];
      int idx = r->head[t] + cur;
      if (idx >= cap) idx -= cap;
      veh_t* e = &r->ring[n->lane_slot0[t] + idx];
      e->xv = pack_xv(x, veh_v(h->xv));
      e->m0 = (h->m0 & ~(63u << 10)) | ((hop + 1) << 10);
      x = veh_x(e->xv);                 /* the tail the next source sees is the stored (rounded) position */
      flag[n->lane_slot0[t] + idx] = 0;
      if (r->trip) r->trip[n->lane_slot0[t] + idx] = r->trip[hslot];
      cur++; tail_x = x; have_tail = 1;
      accepted[src] = 1;
    }
    cnt_add[t] = cur - r->cnt[t];
  }
  /* D: pops, arrivals, refused crossings */
  for (int l = 0; l < L; ++l) {
    if (r->cnt[l] > 0) {
      veh_t* h = veh_at(n, r, l, 0);
      int hslot = (int)(h - r->ring);
      int pop = 0;
      if (flag[hslot] == F_ARRIVE) {
        pop = 1; r->n_arrived++;
        if (r->trip) { /* one tripinfo row */
          if (r->trip_cnt < r->trip_cap) {
            uint32_t t1 = r->trip[hslot];
            uint32_t* row = r->trip_log + 2 * (size_t)r->trip_cnt;
            row[0] = T1_DEPART(t1) | (((uint32_t)(r->cur_sec + 1) & 4095u) << 12) | (M0_ROUTE(h->m0) << 24);
            row[1] = T1_WAIT(t1) | (T1_WCNT(t1) << 16);
          }
          r->trip_cnt++;
        }
      }
      else if (flag[hslot] == F_CROSS) {
        if (accepted[l]) pop = 1;
        else { h->xv = pack_xv(n->lane_len[l] - 0.01f, 0.0f); }
      }
      if (pop) {
        r->head[l]++;
        if (r->head[l] >= n->lane_cap[l]) r->head[l] = 0;
        r->cnt[l]--;
      }
    }
    r->cnt[l] += cnt_add[l];
  }
  /* E: insertion of due vehicles (departPos="random_free" restated on the free tail segment,
   *    large_grid/data/build_file.py:277; at most one insertion per lane per second) */
  if ((int)t_abs < n->horizon)
    for (int q = 0; q < n->n_src; ++q) {
      int due = n->src_due[t_abs * n->n_src + q];
      if (n->src_group && n->n_pint > 0 && due > 0 && n->src_group[q] >= 0) { /* stochastic demand (tsc.h) */
        float ug = u01(rng_draw(key, 0x20000u + (uint32_t)n->src_group[q], 7u));
        int iv = (int)t_abs / n->pint_sec;
        if (iv >= n->n_pint) iv = n->n_pint - 1;
        if (!(ug >= n->src_plo[iv * n->n_src + q] && ug < n->src_phi[iv * n->n_src + q])) due = 0;
      }
      int b = r->backlog[q] + due;
      r->backlog[q] = b > 65535 ? 65535 : b;
    }
  {
    /* evaluate lane ownership on the post-due, pre-insertion backlogs */
    int32_t* owner_ok = (int32_t*)alloca(sizeof(int32_t) * (size_t)n->n_src);
    for (int q = 0; q < n->n_src; ++q) {
      owner_ok[q] = r->backlog[q] > 0;
      if (n->src_group && n->n_pint > 0) {
        /* stochastic demand: the lane goes to the source with the longest backlog (ties: lowest index) */
        for (int p = 0; p < n->n_src; ++p)
          if (p != q && n->src_lane[p] == n->src_lane[q] &&
              (r->backlog[p] > r->backlog[q] || (r->backlog[p] == r->backlog[q] && p < q)))
            owner_ok[q] = 0;
      } else {
        for (int p = 0; p < q; ++p)
          if (n->src_lane[p] == n->src_lane[q] && r->backlog[p] > 0) owner_ok[q] = 0;
      }
    }
    for (int q = 0; q < n->n_src; ++q) {
      if (!owner_ok[q]) continue;
      int lane = n->src_lane[q];
      int cnt = r->cnt[lane];
      if (cnt >= n->lane_cap[lane]) continue;
      float free_back = n->lane_len[lane];
      if (cnt > 0) {
        free_back = veh_x(veh_at(n, r, lane, cnt - 1)->xv) - c->veh_len;
        free_back = free_back - c->min_gap;
      }
      if (free_back < c->veh_len) continue;
      uint32_t qq = (uint32_t)q;
      float u = u01(rng_draw(key, qq, (1u << 16)));
      float pos = c->veh_len + u * (free_back - c->veh_len);
      float su = 0.0f;
      for (uint32_t j = 1; j <= 4; ++j) su = su + u01(rng_draw(key, qq, (1u << 16) | j));
      /* speedFactor ~ N(1, speed_dev) via Irwin-Hall(4): std of sum = sqrt(1/3) */
      float sfr = 1.0f + (c->speed_dev * 1.7320508f) * (su - 2.0f);
      int sfq = (int)((sfr - 0.5f) * 256.0f);
      if (sfq < 0) sfq = 0;
      if (sfq > 255) sfq = 255;
      int cap = n->lane_cap[lane];
      int idx = r->head[lane] + cnt;
      if (idx >= cap) idx -= cap;
      veh_t* e = &r->ring[n->lane_slot0[lane] + idx];
      e->xv = pack_xv(pos, 0.0f);
      e->m0 = ((uint32_t)n->src_route[q] << 16) | ((uint32_t)sfq << 24);
      if (r->trip) r->trip[n->lane_slot0[lane] + idx] = t_abs & 4095u; /* depart second */
      r->cnt[lane] = cnt + 1;
      r->backlog[q]--;
      r->n_departed++;
    }
  }
  r->cur_sec++;
}

/* detector reads at the end of the control step: envs/env.py:325-407 */
static void measure(ref_sim* s, replica_t* r) {
  const tsc_net* n = &s->net;
  const tsc_cfg* c = &s->cfg;
  for (int d = 0; d < n->n_det; ++d) {
    int l = n->det_lane[d];
    float Ll = n->lane_len[l];
    int veh = 0, halt = 0, wait = 0;
    for (int k = 0; k < r->cnt[l]; ++k) {
      veh_t* v = veh_at(n, r, l, k);
      const float vx = veh_x(v->xv);
      if (c->det_len > 0.0f && !(vx > Ll - c->det_len)) break;
      veh++;
      if (veh_v(v->xv) < c->halt_speed) halt++;
      if (k == 0 && vx > 0.0f) wait = (int)M0_WAIT(v->m0);
    }
    r->det_veh[d] = veh; r->det_halt[d] = halt; r->det_wait[d] = wait;
  }
}

static inline float clipf(float x, float hi) {
  if (hi < 0.0f) return x;
  if (x < 0.0f) x = 0.0f;
  if (x > hi) x = hi;
  return x;
}

static void outputs(ref_sim* s, replica_t* r, const float* fp, float* obs, float* reward,
                    float* greward, uint8_t* done) {
  const tsc_net* n = &s->net;
  const tsc_cfg* c = &s->cfg;
  if (obs) {
    for (int k = 0; k < n->n_obs; ++k) {
      int kind = n->obs_kind[k], idx = n->obs_idx[k];
      float v;
      if (kind == 0) v = clipf((float)r->det_veh[idx] / c->norm_wave, c->clip_wave);
      else if (kind == 1) v = clipf((float)r->det_wait[idx] / c->norm_wait, c->clip_wait);
      else v = fp ? fp[idx] : 0.0f;
      obs[k] = n->obs_scale[k] * v;
    }
  }
  float* loc = (float*)alloca(sizeof(float) * (size_t)n->n_nodes);
  float g = 0.0f;
  for (int i = 0; i < n->n_nodes; ++i) {
    int queue = 0, wait = 0;
    for (int d = n->node_det_off[i]; d < n->node_det_off[i + 1]; ++d) {
      int h = r->det_halt[d];
      if (h > c->queue_cap) h = c->queue_cap;
      queue += h; wait += r->det_wait[d];
    }
    float rw;
    if (c->objective == 0) rw = -(float)queue;
    else if (c->objective == 1) rw = -(float)wait;
    else rw = -(float)queue - c->coef_wait * (float)wait;
    loc[i] = rw;
    g = g + rw;
  }
  if (greward) *greward = g;
  if (done) *done = r->cur_sec >= c->episode_length_sec;
  if (reward) {
    int mode = s->train_mode ? c->agent_mode : 0;
    for (int i = 0; i < n->n_nodes; ++i) {
      float rw;
      if (mode == 0) rw = loc[i];
      else if (mode == 1) {
        rw = g;
        if (c->real_net_norm) rw = rw / ((float)n->n_nodes * 20.0f);
      } else {
        rw = loc[i];
        for (int q = n->node_nbr_off[i]; q < n->node_nbr_off[i + 1]; ++q)
          rw = rw + c->coop_gamma * loc[n->node_nbr[q]];
        if (c->real_net_norm)
          rw = rw / ((float)(1 + n->node_nbr_off[i + 1] - n->node_nbr_off[i]) * 20.0f);
      }
      reward[i] = rw;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
static void* dup(const void* p, size_t bytes) {
  void* q = malloc(bytes ? bytes : 1);
  memcpy(q, p, bytes);
  return q;
}
#define DUP(field, count, type) s->net.field = (const type*)dup(net->field, sizeof(type) * (size_t)(count))

ref_sim* ref_create(const tsc_net* net, const tsc_cfg* cfg, int32_t R) {
  ref_sim* s = (ref_sim*)calloc(1, sizeof(ref_sim));
  s->net = *net; s->cfg = *cfg; s->R = R; s->train_mode = 1;
  int L = net->n_lanes, K = net->n_links, N = net->n_nodes;
  DUP(lane_len, L, float); DUP(lane_vmax, L, float); DUP(lane_cap, L, int32_t);
  DUP(lane_slot0, L, int32_t); DUP(lane_inl_off, L + 1, int32_t);
  DUP(lane_inl, net->lane_inl_off[L], int32_t);
  DUP(link_from, K, int32_t); DUP(link_to, K, int32_t); DUP(link_node, K, int32_t);
  DUP(link_tlidx, K, int32_t); DUP(link_vmax, K, float); DUP(link_cross, K, uint32_t);
  DUP(link_merge, K, uint32_t);
  DUP(route_len, net->n_routes, int32_t);
  DUP(route_lane, net->n_routes * net->max_hops, int16_t);
  DUP(route_link, net->n_routes * net->max_hops, int16_t);
  DUP(node_n_phases, N, int32_t); DUP(node_green, N * net->max_phases, uint32_t);
  DUP(node_major, N * net->max_phases, uint32_t);
  DUP(node_det_off, N + 1, int32_t); DUP(det_lane, net->n_det, int32_t);
  DUP(node_nbr_off, N + 1, int32_t); DUP(node_nbr, net->node_nbr_off[N], int32_t);
  DUP(node_obs_off, N + 1, int32_t); DUP(obs_kind, net->n_obs, int32_t);
  DUP(obs_idx, net->n_obs, int32_t); DUP(obs_scale, net->n_obs, float);
  DUP(src_lane, net->n_src, int32_t); DUP(src_route, net->n_src, int32_t);
  DUP(src_due, (size_t)net->horizon * net->n_src, uint8_t);
  if (net->src_group && net->n_pint > 0 && net->pint_sec > 0) {
    DUP(src_group, net->n_src, int32_t);
    DUP(src_plo, (size_t)net->n_pint * net->n_src, float); DUP(src_phi, (size_t)net->n_pint * net->n_src, float);
  } else {
    s->net.src_group = 0; s->net.src_plo = 0; s->net.src_phi = 0; s->net.n_pint = 0; s->net.pint_sec = 1;
  }
  s->rep = (replica_t*)calloc((size_t)R, sizeof(replica_t));
  for (int i = 0; i < R; ++i) {
    replica_t* r = &s->rep[i];
    r->ring = (veh_t*)calloc((size_t)net->n_slots, sizeof(veh_t));
    r->head = (int32_t*)calloc((size_t)L, 4); r->cnt = (int32_t*)calloc((size_t)L, 4);
    r->prev_action = (int32_t*)calloc((size_t)N, 4); r->cur_action = (int32_t*)calloc((size_t)N, 4);
    r->backlog = (int32_t*)calloc((size_t)net->n_src, 4);
    r->det_veh = (int32_t*)calloc((size_t)net->n_det, 4);
    r->det_halt = (int32_t*)calloc((size_t)net->n_det, 4);
    r->det_wait = (int32_t*)calloc((size_t)net->n_det, 4);
  }
  return s;
}

void ref_destroy(ref_sim* s) {
  if (!s) return;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    free(r->ring); free(r->head); free(r->cnt); free(r->prev_action); free(r->cur_action);
    free(r->backlog); free(r->det_veh); free(r->det_halt); free(r->det_wait);
    free(r->trip); free(r->trip_log);
  }
  free(s->rep);
  free(s); /* table copies are leaked on purpose-free builds: test infrastructure */
}

/* reset(): envs/env.py:544-561, _reset_state :444-453 (prev_action = 0) */
void ref_reset(ref_sim* s, const uint64_t* seeds) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    memset(r->head, 0, 4 * (size_t)n->n_lanes); memset(r->cnt, 0, 4 * (size_t)n->n_lanes);
    memset(r->prev_action, 0, 4 * (size_t)n->n_nodes); memset(r->cur_action, 0, 4 * (size_t)n->n_nodes);
    memset(r->backlog, 0, 4 * (size_t)n->n_src);
    memset(r->det_veh, 0, 4 * (size_t)n->n_det); memset(r->det_halt, 0, 4 * (size_t)n->n_det);
    memset(r->det_wait, 0, 4 * (size_t)n->n_det);
    r->cur_sec = 0; r->n_departed = 0; r->n_arrived = 0; r->trip_cnt = 0;
    r->seed_lo = (uint32_t)(seeds[i] & 0xffffffffu); r->seed_hi = (uint32_t)(seeds[i] >> 32);
  }
}

void ref_set_train_mode(ref_sim* s, int32_t m) { s->train_mode = m; }

void ref_observe(ref_sim* s, const float* fp, float* obs) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    measure(s, r);
    outputs(s, r, fp ? fp + (size_t)i * n->n_nodes * n->max_na : 0, obs + (size_t)i * n->n_obs, 0, 0, 0);
  }
}

/* step(action): envs/env.py:566-631.  Replicas are independent; pthreads over replica ranges. */
typedef struct {
  ref_sim* s; int i0, i1;
  const int32_t* action; const float* fp; float* obs; float* reward; float* greward; uint8_t* done;
} step_job;

static void* step_range(void* arg) {
  step_job* j = (step_job*)arg;
  ref_sim* s = j->s;
  const tsc_net* n = &s->net;
  const tsc_cfg* c = &s->cfg;
  float* vnew = (float*)malloc(4 * (size_t)n->n_slots);
  float* xnew = (float*)malloc(4 * (size_t)n->n_slots);
  uint8_t* flag = (uint8_t*)calloc((size_t)n->n_slots, 1);
  uint8_t* hblk = (uint8_t*)calloc((size_t)n->n_lanes, 1);
  float* head_lim = (float*)malloc(4 * (size_t)n->n_lanes);
  uint32_t* approach = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* open = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* major = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* ymask = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint8_t* accepted = (uint8_t*)malloc((size_t)n->n_lanes);
  int32_t* cnt_add = (int32_t*)malloc(4 * (size_t)n->n_lanes);
  for (int i = j->i0; i < j->i1; ++i) {
    replica_t* r = &s->rep[i];
    for (int k = 0; k < n->n_nodes; ++k) r->cur_action[k] = j->action[(size_t)i * n->n_nodes + k];
    for (int t = 0; t < c->control_interval_sec; ++t)
      substep(s, r, t < c->yellow_interval_sec, vnew, xnew, flag, hblk, head_lim, approach, open, major,
              ymask, accepted, cnt_add);
    /* node.prev_action = action (set in the 'yellow' call, envs/env.py:134) */
    for (int k = 0; k < n->n_nodes; ++k) r->prev_action[k] = r->cur_action[k];
    measure(s, r);
    outputs(s, r, j->fp ? j->fp + (size_t)i * n->n_nodes * n->max_na : 0,
            j->obs ? j->obs + (size_t)i * n->n_obs : 0,
            j->reward ? j->reward + (size_t)i * n->n_nodes : 0,
            j->greward ? j->greward + i : 0, j->done ? j->done + i : 0);
  }
  free(vnew); free(xnew); free(flag); free(hblk); free(head_lim); free(approach); free(open); free(major);
  free(ymask); free(accepted); free(cnt_add);
  return 0;
}

void ref_step_mt(ref_sim* s, const int32_t* action, const float* fp, float* obs, float* reward,
                 float* greward, uint8_t* done, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > s->R) n_threads = s->R;
  if (n_threads > 256) n_threads = 256;
  step_job jobs[256];
  pthread_t th[256];
  int per = (s->R + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    int i0 = t * per, i1 = i0 + per > s->R ? s->R : i0 + per;
    if (i0 > s->R) i0 = s->R;
    step_job jb = {s, i0, i1, action, fp, obs, reward, greward, done};
    jobs[t] = jb;
  }
  if (n_threads == 1) { step_range(&jobs[0]); return; }
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, step_range, &jobs[t]);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
}

/* n_steps control steps in one call: every thread advances its replica range through ALL steps (replicas are
 * independent), so threads are created once — the fair way to time the CPU path on many cores.
 * actions [n_act][R][n_nodes] are cycled; obs / reward receive the last step's outputs. */
typedef struct { step_job job; const int32_t* actions; int32_t n_act, n_steps; } run_job;

static void* run_range(void* arg) {
  run_job* rj = (run_job*)arg;
  const tsc_net* n = &rj->job.s->net;
  const size_t stride = (size_t)rj->job.s->R * n->n_nodes;
  for (int t = 0; t < rj->n_steps; ++t) {
    step_job j = rj->job;
    j.action = rj->actions + (size_t)(t % rj->n_act) * stride;
    step_range(&j);
  }
  return 0;
}

void ref_run_mt(ref_sim* s, const int32_t* actions, int32_t n_act, int32_t n_steps, const float* fp, float* obs,
                float* reward, float* greward, uint8_t* done, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > s->R) n_threads = s->R;
  if (n_threads > 256) n_threads = 256;
  run_job jobs[256];
  pthread_t th[256];
  int per = (s->R + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    int i0 = t * per, i1 = i0 + per > s->R ? s->R : i0 + per;
    if (i0 > s->R) i0 = s->R;
    step_job jb = {s, i0, i1, 0, fp, obs, reward, greward, done};
    jobs[t].job = jb; jobs[t].actions = actions; jobs[t].n_act = n_act; jobs[t].n_steps = n_steps;
  }
  if (n_threads == 1) { run_range(&jobs[0]); return; }
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, run_range, &jobs[t]);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
}

void ref_step(ref_sim* s, const int32_t* action, const float* fp, float* obs, float* reward,
              float* greward, uint8_t* done) {
  ref_step_mt(s, action, fp, obs, reward, greward, done, 1);
}

/* ---- evaluation / recording path (envs/env.py:409-437, 498-542) --------------------------------- */
/* record mode on: allocate the trip words and the arrival log (call right after ref_reset) */
void ref_set_record(ref_sim* s, int32_t on) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    free(r->trip); free(r->trip_log);
    r->trip = 0; r->trip_log = 0; r->trip_cnt = 0; r->trip_cap = 0;
    if (on) {
      r->trip = (uint32_t*)calloc((size_t)n->n_slots, 4);
      r->trip_cap = 8192;
      r->trip_log = (uint32_t*)calloc((size_t)r->trip_cap * 2, 4);
    }
  }
}

/* _measure_traffic_step for every replica: out[r] = {n_live, departed_total, arrived_total, avg_wait, avg_speed,
 * avg_queue, std_queue, backlog}; queue = vehicles slower than 0.1 m/s on the whole detector lane
 * (lane.getLastStepHaltingNumber, envs/env.py:425) */
void ref_traffic_stats(ref_sim* s, float* out) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    int V = 0, wsum = 0, backlog = 0;
    float sp = 0.0f;
    int32_t* halt = (int32_t*)calloc((size_t)n->n_lanes, 4);
    for (int l = 0; l < n->n_lanes; ++l)
      for (int k = 0; k < r->cnt[l]; ++k) {
        veh_t* v = veh_at(n, r, l, k);
        V++; wsum += (int)M0_WAIT(v->m0); sp = sp + veh_v(v->xv);
        if (veh_v(v->xv) < 0.1f) halt[l]++;
      }
    float q = 0.0f, q2 = 0.0f;
    for (int d = 0; d < n->n_det; ++d) { float h = (float)halt[n->det_lane[d]]; q = q + h; q2 = q2 + h * h; }
    float nd = (float)(n->n_det > 0 ? n->n_det : 1);
    float mq = q / nd;
    float var = q2 / nd - mq * mq;
    if (var < 0.0f) var = 0.0f;
    for (int q3 = 0; q3 < n->n_src; ++q3) backlog += r->backlog[q3];
    float* o = out + 8 * (size_t)i;
    o[0] = (float)V; o[1] = (float)r->n_departed; o[2] = (float)r->n_arrived;
    o[3] = V > 0 ? (float)wsum / (float)V : 0.0f; o[4] = V > 0 ? sp / (float)V : 0.0f;
    o[5] = mq; o[6] = sqrtf(var); o[7] = (float)backlog;
    free(halt);
  }
}

/* step(action) one simulated second at a time, with the traffic statistics after every second:
 * sub_stats [R][control_interval_sec][8].  Same results as ref_step. */
void ref_step_record(ref_sim* s, const int32_t* action, const float* fp, float* obs, float* reward,
                     float* greward, uint8_t* done, float* sub_stats) {
  const tsc_net* n = &s->net;
  const tsc_cfg* c = &s->cfg;
  float* vnew = (float*)malloc(4 * (size_t)n->n_slots);
  float* xnew = (float*)malloc(4 * (size_t)n->n_slots);
  uint8_t* flag = (uint8_t*)calloc((size_t)n->n_slots, 1);
  uint8_t* hblk = (uint8_t*)calloc((size_t)n->n_lanes, 1);
  float* head_lim = (float*)malloc(4 * (size_t)n->n_lanes);
  uint32_t* approach = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* open = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* major = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint32_t* ymask = (uint32_t*)malloc(4 * (size_t)n->n_nodes);
  uint8_t* accepted = (uint8_t*)malloc((size_t)n->n_lanes);
  int32_t* cnt_add = (int32_t*)malloc(4 * (size_t)n->n_lanes);
  float* st = (float*)malloc(4 * 8 * (size_t)s->R);
  const int ci = c->control_interval_sec;
  for (int i = 0; i < s->R; ++i)
    for (int k = 0; k < n->n_nodes; ++k) s->rep[i].cur_action[k] = action[(size_t)i * n->n_nodes + k];
  for (int t = 0; t < ci; ++t) {
    for (int i = 0; i < s->R; ++i)
      substep(s, &s->rep[i], t < c->yellow_interval_sec, vnew, xnew, flag, hblk, head_lim, approach, open, major, ymask,
              accepted, cnt_add);
    if (sub_stats) {
      ref_traffic_stats(s, st);
      for (int i = 0; i < s->R; ++i) memcpy(sub_stats + ((size_t)i * ci + t) * 8, st + 8 * (size_t)i, 32);
    }
  }
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    for (int k = 0; k < n->n_nodes; ++k) r->prev_action[k] = r->cur_action[k];
    measure(s, r);
    outputs(s, r, fp ? fp + (size_t)i * n->n_nodes * n->max_na : 0, obs ? obs + (size_t)i * n->n_obs : 0,
            reward ? reward + (size_t)i * n->n_nodes : 0, greward ? greward + i : 0, done ? done + i : 0);
  }
  free(vnew); free(xnew); free(flag); free(hblk); free(head_lim); free(approach); free(open); free(major);
  free(ymask); free(accepted); free(cnt_add); free(st);
}

/* tripinfo rows of one replica (arrival order): rows [n][2] as in replica_t.trip_log */
void ref_get_trips(ref_sim* s, int32_t replica, uint32_t* rows, int32_t max_rows, int32_t* n_rows) {
  replica_t* r = &s->rep[replica];
  int m = r->trip_cnt < r->trip_cap ? r->trip_cnt : r->trip_cap;
  if (m > max_rows) m = max_rows;
  if (m > 0 && r->trip_log) memcpy(rows, r->trip_log, 8 * (size_t)m);
  *n_rows = r->trip_log ? m : 0;
}

void ref_get_counts(ref_sim* s, int32_t* veh, int32_t* halt, int32_t* headwait, int32_t* phase) {
  const tsc_net* n = &s->net;
  for (int i = 0; i < s->R; ++i) {
    replica_t* r = &s->rep[i];
    if (veh) memcpy(veh + (size_t)i * n->n_det, r->det_veh, 4 * (size_t)n->n_det);
    if (halt) memcpy(halt + (size_t)i * n->n_det, r->det_halt, 4 * (size_t)n->n_det);
    if (headwait) memcpy(headwait + (size_t)i * n->n_det, r->det_wait, 4 * (size_t)n->n_det);
    if (phase) memcpy(phase + (size_t)i * n->n_nodes, r->prev_action, 4 * (size_t)n->n_nodes);
  }
}

/* canonical state dump, same format as tsc_dump_state */
void ref_dump_state(ref_sim* s, int32_t replica, int32_t* lane_cnt, uint32_t* veh, int32_t* n_veh) {
  const tsc_net* n = &s->net;
  replica_t* r = &s->rep[replica];
  int w = 0;
  for (int l = 0; l < n->n_lanes; ++l) {
    lane_cnt[l] = r->cnt[l];
    for (int k = 0; k < r->cnt[l]; ++k) {
      veh_t* v = veh_at(n, r, l, k);
      float fx = veh_x(v->xv), fv = veh_v(v->xv);
      memcpy(veh + 3 * (size_t)w, &fx, 4); memcpy(veh + 3 * (size_t)w + 1, &fv, 4); veh[3 * (size_t)w + 2] = v->m0;
      w++;
    }
  }
  *n_veh = w;
}

/* test probe of the car-following helpers: kind 0 brake_gap(v=a, b=b), 1 stop_speed(gap=a, b=b, tau=c),
 * 2 follow_speed(gap=a, v_lead=b, b=c, tau=d), 3 free_speed(dist=a, target=b, b=c) */
float ref_probe_krauss(int32_t kind, float a, float b, float c, float d) {
  switch (kind) {
    case 0: return brake_gap(a, b, 1.0f / b);
    case 1: return stop_speed(a, b, 1.0f / b, c);
    case 2: return follow_speed(a, b, c, 1.0f / c, d);
    default: return free_speed(a, b, c);
  }
}

/* pending (not yet inserted) vehicles of every demand source of one replica */
void ref_get_backlog(ref_sim* s, int32_t replica, int32_t* out /* [n_src] */) {
  memcpy(out, s->rep[replica].backlog, 4 * (size_t)s->net.n_src);
}

void ref_get_misc(ref_sim* s, int32_t replica, int32_t* out /* cur_sec, departed, arrived, backlog_sum, live */) {
  const tsc_net* n = &s->net;
  replica_t* r = &s->rep[replica];
  int b = 0, live = 0;
  for (int q = 0; q < n->n_src; ++q) b += r->backlog[q];
  for (int l = 0; l < n->n_lanes; ++l) live += r->cnt[l];
  out[0] = r->cur_sec; out[1] = r->n_departed; out[2] = r->n_arrived; out[3] = b; out[4] = live;
}
