"""`BatchedA2C`: all A per-intersection actor-critics, R replicas at once, on one GPU.

Device-side counterpart of reference agents/models.py (IA2C/MA2C) + agents/policies.py
(LstmACPolicy / FPLstmACPolicy) + agents/utils.py (OnPolicyBuffer).  Semantics kept:
  * two separate LSTM networks per agent, state zeroed inside the cell on a pre-decision done
    (agents/utils.py:104-105); 'v'-only forward does not advance the state (agents/policies.py:127-135);
  * BPTT over n_step from `states_bw`, refreshed from `states_fw` after each update (:153);
  * loss of agents/policies.py:41-52, per-agent clip_by_global_norm, TF1 RMSProp (:54-61);
  * n-step returns of OnPolicyBuffer (agents/utils.py:202-214), reward /= reward_norm then clip
    (agents/models.py:222-229).
Replicas share the weights: the gradient is the mean over replicas (and over ranks: one
`all_reduce(SUM)` of the flat gradient per update, then identical updates everywhere).

Shipping path (`use_tc`, the default): every kernel is hand-written — the fused tcgen05 forward
(csrc/tsc_policy_tc.cu: fc front end, gate GEMM, LSTM cell, heads, sampling, bf16 activation store), the tcgen05
update (BPTT with TMA operand copies, dX = dZ.Wx^T, LSTM and fc weight gradients) and the SIMT kernels of
csrc/tsc_learn.cu (loss / head gradients, returns, clip + RMSProp).  No library GEMM runs on it.
`use_tc=False` selects the plain fp32 twin kernels (fc_embed, lstm_seq_fwd / bwd, heads, fc_bwd) that the reference
goldens pin at 2e-4; only on that path do the three plain batched products (X.Wx, dZ.Wx^T, [X|H]^T.dZ) go through
`torch.baddbmm / bmm` (fp32, TF32 only when `allow_tf32`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .. import dist as _dist
from .layout import PolicyLayout


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedA2C:
    def __init__(self, layout: PolicyLayout, n_replicas: int, n_step: int, gamma: float = 0.99,
                 v_coef: float = 0.5, max_grad_norm: float = 40.0, alpha: float = 0.99, eps: float = 1e-5,
                 reward_norm: float = 1.0, reward_clip: float = 0.0, seed: int = 0, device: int = 0,
                 chunk: int = 1024, replica0: int = 0, total_replicas: Optional[int] = None,
                 process_group=None, allow_tf32: bool = True, use_tc: bool = True,
                 store_acts: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedA2C needs a CUDA device (no CPU fallback exists)")
        self.lay, self.R, self.T = layout, int(n_replicas), int(n_step)
        self.gamma, self.v_coef, self.max_grad_norm = gamma, v_coef, max_grad_norm
        self.alpha, self.eps = alpha, eps
        self.reward_norm, self.reward_clip = reward_norm, reward_clip
        self.seed, self.replica0 = int(seed), int(replica0)
        self.total_replicas = int(total_replicas or n_replicas)
        self.pg = process_group
        self.allow_tf32 = allow_tf32
        self.dev = torch.device("cuda", device)
        self.chunk = min(int(chunk), self.R)
        lib = _lib.lib()
        self._cd = layout.as_c()
        h = C.c_void_p()
        _lib.check(lib.tscl_create(C.byref(self._cd), C.c_int32(device), C.byref(h)))
        self._h = h
        L, R, T, U, A = layout, self.R, self.T, layout.U, layout.A
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.P = torch.from_numpy(layout.init_params(seed)).to(self.dev)
        self.G = torch.zeros_like(self.P)
        self.MS = torch.ones_like(self.P)                      # TF1 RMSProp slot "rms" starts at 1
        self.agent_of = torch.from_numpy(layout.agent_of).to(self.dev)
        self.norms = torch.zeros(A, **f32)
        self.stats = torch.zeros(4, **f32)
        self.pv, self.gv = layout.views(self.P), layout.views(self.G)
        # recurrent state: [U][R][h] each
        self.c_fw = torch.zeros(U, R, L.h, **f32); self.h_fw = torch.zeros(U, R, L.h, **f32)
        self.c_bw = torch.zeros_like(self.c_fw); self.h_bw = torch.zeros_like(self.h_fw)
        self.c_tmp = torch.zeros_like(self.c_fw); self.h_tmp = torch.zeros_like(self.h_fw)
        # per-step work buffers
        self.X1 = torch.empty(U, R, L.dx, **f32)
        self.Z1 = torch.empty(U, R, 4 * L.h, **f32)
        self.H1 = torch.empty(U, R, L.h, **f32)
        self.pi = torch.zeros(R, A, L.max_na, **f32)
        self.val = torch.zeros(R, A, **f32)
        self.act = torch.zeros(R, A, dtype=torch.int32, device=self.dev)
        self.boot = torch.zeros(R, A, **f32)
        # rollout storage; obs slot t is what forward consumed at step t, slot T is the next obs
        self.obs_hist = torch.zeros(T + 1, R, L.n_obs, **f32)
        self.act_hist = torch.zeros(T, R, A, dtype=torch.int32, device=self.dev)
        self.rew_hist = torch.zeros(T, R, A, **f32)
        self.val_hist = torch.zeros(T, R, A, **f32)
        self.Rs = torch.zeros(T, R, A, **f32); self.Adv = torch.zeros(T, R, A, **f32)
        self.done_pre = [0.0] * T
        self.done_post = [0.0] * T
        self.last_done = False      # OnPolicyBuffer.reset(done) carries the last done (agents/utils.py:187-193)
        self.t = 0
        self.n_forward = 0
        self._one = torch.zeros(1, **f32)
        self._upd_bufs = None
        self.kernel_launches = 0
        # fused tensor-core forward (csrc/tsc_policy_tc.cu): bf16 image of [Wx;Wh], refreshed after every update
        self.use_tc = bool(use_tc) and (L.dx % 16 == 0) and layout.kw > 0    # else: the fp32 kernels
        self.tc_v2 = self.use_tc and (L.dx % 32 == 0)          # fc front end on the tensor cores too
        self.Wp = torch.zeros(U, ((L.dx + L.h) // 8) * 4 * L.h * 8 + 8 * L.dx * 8, dtype=torch.bfloat16,
                              device=self.dev)
        self.Wt = torch.zeros(U, 32, L.h, 8, dtype=torch.bfloat16, device=self.dev)    # Wh^T image for the BPTT MMA
        self.Wxt = torch.zeros(U, 32, L.dx, 8, dtype=torch.bfloat16, device=self.dev)  # Wx^T image: dX fused into the BPTT
        # measured (R = 8192, 1 x B200): fusing dX into the BPTT step lengthens the serial per-step chain (update 72.5 ->
        # 92.5 ms), so the default keeps dX as a separate product; `dx_fused = True` selects the fused kernel (tested)
        self.dx_fused = False
        self.dx_fusable = self.use_tc and L.dx % 32 == 0 and L.dx <= 256
        # stand-alone dX = dZ . Wx^T kernel (tscl_dx_tc); False falls back to the library GEMM (A/B measurements only)
        self.dx_own = self.use_tc and L.dx % 16 == 0 and L.dx <= 224 and os.environ.get("TSC_DX_LIBRARY", "0") != "1"
        self.bwd_tc = self.use_tc
        self.fc_bwd_tc = self.use_tc and layout.fc_bwd_tc_ok     # front-end weight gradients on the tensor cores
        self.wgrad_tc = self.use_tc and L.dx % 8 == 0 and L.dx <= 240   # LSTM weight gradients on the tensor cores
        self.fused_heads = True                                  # head weight gradients inside tscl_heads_loss
        self.pack_weights()
        # bf16 activation store of the rollout's own forward pass (written by the v2 kernel): the update then
        # back-propagates through it instead of recomputing fc + gate GEMM + LSTM forward.
        need = U * T * R * (L.dx + 4 * L.h + 2 * L.h) * 2
        if store_acts is None:
            free, _ = torch.cuda.mem_get_info(self.dev)
            store_acts = self.tc_v2 and need < 0.5 * free
        self.store_acts = bool(store_acts) and self.tc_v2 and (R % self.chunk == 0)
        self.st_x = self.st_g = self.st_c = self.st_h = None
        if self.store_acts:       # [R/chunk][U][T][chunk][w]: every update chunk is one contiguous block
            bf = dict(dtype=torch.bfloat16, device=self.dev)
            nc, rc_ = R // self.chunk, self.chunk
            self.st_x = torch.zeros(nc, U, T, rc_, L.dx, **bf); self.st_g = torch.zeros(nc, U, T, rc_, 4 * L.h, **bf)
            self.st_c = torch.zeros(nc, U, T, rc_, L.h, **bf); self.st_h = torch.zeros(nc, U, T, rc_, L.h, **bf)
        self._acts_ok = [False] * T      # step t of the current rollout was produced by a storing forward()

    def close(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().tscl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def pack_weights(self):
        if self.use_tc:
            _lib.check(_lib.lib().tscl_pack_weights(self._h, _p(self.P), _p(self.Wp), self._st()))
            _lib.check(_lib.lib().tscl_pack_wht(self._h, _p(self.P), _p(self.Wt), self._st()))
            if self.dx_fusable or self.dx_own:
                _lib.check(_lib.lib().tscl_pack_wxt(self._h, _p(self.P), _p(self.Wxt), self._st()))
            self.wx_b = self.pv["wx"].to(torch.bfloat16)       # operand of the separate product dX = dZ . Wx^T
            self.kernel_launches += 3

    def _mm(self):
        # cuBLAS fp32 (or TF32 when allowed) for the plain batched GEMMs
        torch.backends.cuda.matmul.allow_tf32 = bool(self.allow_tf32)

    # ------------------------------------------------------------------------------------------
    def reset(self):
        """model.reset(): zero the LSTM states (agents/models.py:218-220, policies.py:120-123)."""
        for t in (self.c_fw, self.h_fw, self.c_bw, self.h_bw):
            t.zero_()

    def forward(self, obs: torch.Tensor, done: bool, out_type: str = "pv", sample: bool = True, to_hist: bool = False):
        """One decision for all replicas/agents.  obs [R, n_obs] device tensor.  Returns
        (pi [R, A, max_na], val [R, A], act [R, A] or None); 'v' does not advance the state.
        `to_hist` (fused tensor-core forward only): values / actions are written straight into the rollout slots
        val_hist[t] / act_hist[t] (what add_transition would copy there); those views are returned."""
        L, R, lib = self.lay, self.R, _lib.lib()
        commit = "p" in out_type
        want_act = sample and commit
        to_hist = to_hist and self.use_tc and self.tc_v2 and want_act and self.t < self.T
        val_o, act_o = (self.val_hist[self.t], self.act_hist[self.t]) if to_hist else (self.val, self.act)
        if self.use_tc:
            c1, h1 = (self.c_fw, self.h_fw) if commit else (self.c_tmp, self.h_tmp)
            args = (self._h, _p(self.P), _p(self.Wp), _p(obs), C.c_int64(R), _p(self.c_fw), _p(self.h_fw), _p(c1),
                    _p(h1), _p(self.pi), _p(val_o), _p(act_o) if want_act else None,
                    C.c_int32(1 if done else 0), C.c_uint64(self.seed), C.c_int64(self.n_forward),
                    C.c_int64(self.replica0), None)
            if self.tc_v2:
                store = self.store_acts and commit and self.t < self.T
                st = (_p(self.st_x), _p(self.st_g), _p(self.st_c), _p(self.st_h)) if store else (None,) * 4
                _lib.check(lib.tscl_policy_step_v2(*args, *st, C.c_int32(self.t if store else 0), C.c_int32(self.T),
                                                   C.c_int64(self.chunk), self._st()))
                if commit and self.t < self.T:
                    self._acts_ok[self.t] = store
            else:
                _lib.check(lib.tscl_policy_step(*args, C.c_int32(0), self._st()))
            self.kernel_launches += 1
            if commit:
                self.n_forward += 1
            self._hist_direct = to_hist
            return self.pi, val_o, (act_o if want_act else None)
        self._mm()
        dflag = self._one.fill_(1.0 if done else 0.0)
        _lib.check(lib.tscl_fc_embed(self._h, _p(self.P), _p(obs), C.c_int64(R), C.c_int64(R), C.c_int64(0),
                                     _p(self.X1), self._st()))
        torch.baddbmm(self.pv["bl"].unsqueeze(1), self.X1, self.pv["wx"], out=self.Z1)
        c1, h1 = (self.c_fw, self.h_fw) if commit else (self.c_tmp, self.h_tmp)
        _lib.check(lib.tscl_lstm_seq_fwd(self._h, _p(self.P), _p(self.Z1), None, _p(self.H1), None, _p(self.c_fw),
                                         _p(self.h_fw), _p(c1), _p(h1), _p(dflag), C.c_int32(1), C.c_int64(R),
                                         C.c_int64(R), C.c_int64(0), self._st()))
        _lib.check(lib.tscl_heads(self._h, _p(self.P), _p(self.H1), C.c_int64(R), _p(self.pi), _p(self.val),
                                  _p(self.act) if want_act else None, C.c_uint64(self.seed),
                                  C.c_int64(self.n_forward), C.c_int64(self.replica0), self._st()))
        self.kernel_launches += 3
        if commit:
            self.n_forward += 1
        return self.pi, self.val, (self.act if want_act else None)

    def forward_range(self, r0: int, n: int, done: bool, t: int, n_forward: int, stream=None, to_hist: bool = False):
        """forward() for the replica range [r0, r0 + n) only, on the current stream, for rollout slot `t` and decision
        counter `n_forward` (the caller may run ranges one step apart): reads obs_slot(t)[r0:r0+n], advances that
        range's recurrent state and writes its slices of pi / val / act (and of the activation store).  Ranges are
        independent.  Tensor-core path only; bookkeeping of the step: end_forward_ranges().
        `stream`: raw CUDA stream handle (default: torch's current stream).  `to_hist`: actions / values go straight into
        the rollout slots act_hist[t] / val_hist[t] (what add_transition would copy there) and those views are returned."""
        assert self.use_tc and self.tc_v2, "forward_range needs the fused tensor-core forward"
        L, R, A = self.lay, self.R, self.lay.A
        store = self.store_acts and t < self.T
        st = (_p(self.st_x), _p(self.st_g), _p(self.st_c), _p(self.st_h)) if store else (None,) * 4
        off = lambda t_, per_row: C.c_void_p(t_.data_ptr() + r0 * per_row * t_.element_size())
        _lib.check(_lib.lib().tscl_policy_step_v2r(
            self._h, _p(self.P), _p(self.Wp), off(self.obs_hist[t], L.n_obs), C.c_int64(n),
            off(self.c_fw, L.h), off(self.h_fw, L.h), off(self.c_fw, L.h), off(self.h_fw, L.h),
            off(self.pi, A * L.max_na), off(self.val_hist[t] if to_hist else self.val, A),
            off(self.act_hist[t] if to_hist else self.act, A), C.c_int32(1 if done else 0),
            C.c_uint64(self.seed), C.c_int64(n_forward), C.c_int64(self.replica0 + r0), None, *st,
            C.c_int32(t if store else 0), C.c_int32(self.T), C.c_int64(self.chunk), C.c_int64(R), C.c_int64(r0),
            self._st() if stream is None else stream))
        self.kernel_launches += 1
        if t < self.T:
            self._acts_ok[t] = store
        if to_hist:
            return self.pi[r0:r0 + n], self.val_hist[t, r0:r0 + n], self.act_hist[t, r0:r0 + n]
        return self.pi[r0:r0 + n], self.val[r0:r0 + n], self.act[r0:r0 + n]

    def end_forward_ranges(self):
        """One decision step has been issued for every range."""
        self.n_forward += 1

    def add_transition_range(self, r0: int, n: int, reward: torch.Tensor):
        """add_transition() data movement for one replica range (reward [n, A]); finish the step with
        end_transition_ranges(done_pre, done_post)."""
        t = self.t
        r = reward
        if self.reward_norm:
            r = r / self.reward_norm
        if self.reward_clip:
            r = torch.clamp(r, -self.reward_clip, self.reward_clip)
        self.rew_hist[t, r0:r0 + n].copy_(r)
        self.act_hist[t, r0:r0 + n].copy_(self.act[r0:r0 + n])
        self.val_hist[t, r0:r0 + n].copy_(self.val[r0:r0 + n])

    def end_transition_ranges(self, done_pre: bool, done_post: bool):
        t = self.t
        self.done_pre[t] = 1.0 if done_pre else 0.0
        self.done_post[t] = 1.0 if done_post else 0.0
        self.t += 1

    # ------------------------------------------------------------------------------------------
    def obs_slot(self, t: Optional[int] = None) -> torch.Tensor:
        return self.obs_hist[self.t if t is None else t]

    def add_transition(self, reward: torch.Tensor, done_pre: bool, done_post: bool,
                       act: Optional[torch.Tensor] = None, val: Optional[torch.Tensor] = None):
        """Record step t: obs must already be in obs_slot(t) (the env writes there); actions and
        values default to the ones of the last forward().  agents/models.py:222-229."""
        t = self.t
        r = reward
        if self.reward_norm:
            r = r / self.reward_norm
        if self.reward_clip:
            r = torch.clamp(r, -self.reward_clip, self.reward_clip)
        self.rew_hist[t].copy_(r)
        self.act_hist[t].copy_(self.act if act is None else act)
        self.val_hist[t].copy_(self.val if val is None else val)
        self.done_pre[t] = 1.0 if done_pre else 0.0
        self.done_post[t] = 1.0 if done_post else 0.0
        self.t += 1

    def add_transition_device(self, reward: torch.Tensor, greward: torch.Tensor, rew_acc: torch.Tensor, done_pre: bool,
                              done_post: bool):
        """add_transition() of the device-resident loop in ONE launch: normalised / clipped reward into the rollout slot
        and the episode sum of the global reward; actions / values are already in their slots when the last forward ran
        with to_hist=True (copied otherwise).  Same arithmetic as add_transition (r * (1 / norm), clamp)."""
        t = self.t
        _lib.check(_lib.lib().tscl_device_transition(
            self._h, _p(reward), _p(self.rew_hist[t]), C.c_int64(reward.numel()), C.c_float(self.reward_norm or 0.0),
            C.c_float(self.reward_clip or 0.0), _p(greward), _p(rew_acc), C.c_int64(greward.numel()), self._st()))
        if not getattr(self, "_hist_direct", False):
            self.act_hist[t].copy_(self.act)
            self.val_hist[t].copy_(self.val)
        self.done_pre[t] = 1.0 if done_pre else 0.0
        self.done_post[t] = 1.0 if done_post else 0.0
        self.t += 1

    def _bufs(self, rc, lean=False):
        """Work buffers of one update chunk.  `lean`: every consumer reads the bf16 activation store itself and dZ / dX
        travel as bf16 between the tensor-core kernels, so only dH (fp32), dZb and dXb (bf16) exist; the fp32 set is
        allocated the first time a fallback path needs it."""
        L, T, U = self.lay, self.T, self.lay.U
        f32 = dict(dtype=torch.float32, device=self.dev)
        if self._upd_bufs is None or self._upd_bufs["rc"] < rc:
            M = T * rc
            self._upd_bufs = dict(rc=rc, dH=torch.empty(U, M, L.h, **f32))
        b = self._upd_bufs
        M = T * b["rc"]
        if lean and "dZb" not in b:
            b.update(dZb=torch.empty(U, M, 4 * L.h, dtype=torch.bfloat16, device=self.dev),
                     dXb=torch.empty(U, M, L.dx, dtype=torch.bfloat16, device=self.dev))
        if not lean and "X" not in b:
            b.update(ZG=torch.empty(U, M, 4 * L.h, **f32), dX=torch.empty(U, M, L.dx, **f32),
                     X=torch.empty(U, M, L.dx, **f32), C=torch.empty(U, M, L.h, **f32), H=torch.empty(U, M, L.h, **f32),
                     Hp=torch.empty(U, M, L.h, **f32), dlog=torch.empty(U, M, L.max_na, **f32))
        return b

    def backward(self, boot: Optional[torch.Tensor], lr: float, beta: float):
        """One A2C update from the stored n_step rollout (agents/models.py:174-183).  `boot` is the
        bootstrap value [R, A] (None / zeros when the episode ended, utils.py:186-190)."""
        assert self.t == self.T, "rollout buffer not full"
        L, R, T, U, A, lib = self.lay, self.R, self.T, self.lay.U, self.lay.A, _lib.lib()
        self._mm()
        st = self._st
        f32 = dict(dtype=torch.float32, device=self.dev)
        dpre = torch.tensor(self.done_pre, **f32)
        dpost = torch.tensor(self.done_post, **f32)
        if boot is None:
            self.boot.zero_()
        else:
            self.boot.copy_(boot)
        _lib.check(lib.tscl_returns(self._h, _p(self.rew_hist), _p(self.val_hist), _p(self.boot), _p(dpost),
                                    C.c_float(self.gamma), C.c_int32(T), C.c_int64(R), _p(self.Rs), _p(self.Adv), st()))
        self.G.zero_()
        self.stats.zero_()
        scale = _dist.grad_scale(T, 1, self.total_replicas)       # local SUM x 1/(n_step * R_total); ranks add up
        use_store = self.store_acts and all(self._acts_ok)
        n_obs = L.n_obs
        for r0 in range(0, R, self.chunk):
            rc = min(self.chunk, R - r0)
            M = T * rc
            ci = r0 // self.chunk
            all_tc = use_store and self.bwd_tc and self.fc_bwd_tc and self.wgrad_tc and self.fused_heads
            b = self._bufs(rc, lean=all_tc)
            X = Cc = H = Hp = dlog = ZG = dX = dZb = dXb = None
            bf16 = dict(dtype=torch.bfloat16, device=self.dev)
            if b["rc"] == rc:
                dH = b["dH"]
                if all_tc:
                    dZb, dXb = b["dZb"], b["dXb"]
                else:
                    ZG, dX, X, Cc, H, Hp, dlog = (b[k] for k in ("ZG", "dX", "X", "C", "H", "Hp", "dlog"))
            else:               # tail chunk: dense temporaries of the right shape
                dH = torch.empty(U, M, L.h, **f32)
                if all_tc:
                    dZb, dXb = torch.empty(U, M, 4 * L.h, **bf16), torch.empty(U, M, L.dx, **bf16)
                else:
                    ZG, dX, X, Cc, H, Hp, dlog = (torch.empty(U, M, s_, **f32) for s_ in
                                                  (4 * L.h, L.dx, L.dx, L.h, L.h, L.h, L.max_na))
            obs0 = self.obs_hist[0, r0:]
            if all_tc:
                pass        # every consumer reads the bf16 activation store itself
            elif use_store:
                # activations of the rollout's forward pass (bf16 store -> fp32 chunk buffers)
                direct = self.bwd_tc        # the tensor-core BPTT kernel reads gates / c from the store itself
                _lib.check(lib.tscl_unpack_store(self._h, _p(self.st_x[ci]), _p(self.st_g[ci]), _p(self.st_c[ci]),
                                                 _p(self.st_h[ci]), _p(X), None if direct else _p(ZG),
                                                 None if direct else _p(Cc), _p(H), _p(Hp), _p(self.h_bw),
                                                 _p(dpre), C.c_int32(T), C.c_int64(rc), C.c_int64(R), C.c_int64(r0),
                                                 st()))
            else:
                _lib.check(lib.tscl_fc_embed(self._h, _p(self.P), _p(obs0), C.c_int64(M), C.c_int64(rc),
                                             C.c_int64(R * n_obs), _p(X), st()))
                torch.baddbmm(self.pv["bl"].unsqueeze(1), X, self.pv["wx"], out=ZG)
                _lib.check(lib.tscl_lstm_seq_fwd(self._h, _p(self.P), _p(ZG), _p(Cc), _p(H), _p(Hp), _p(self.c_bw),
                                                 _p(self.h_bw), None, None, _p(dpre), C.c_int32(T), C.c_int64(rc),
                                                 C.c_int64(R), C.c_int64(r0), st()))
            hb = _p(self.st_h[ci]) if use_store else None
            _lib.check(lib.tscl_heads_loss(self._h, _p(self.P), None if all_tc else _p(H), _p(self.act_hist[0, r0:]),
                                           _p(self.Rs[0, r0:]), _p(self.Adv[0, r0:]), C.c_int64(M), C.c_int64(rc),
                                           C.c_int64(R * A), C.c_float(self.v_coef), C.c_float(beta), C.c_float(scale),
                                           None if self.fused_heads else _p(dlog), _p(dH), _p(self.stats),
                                           hb if all_tc else None, _p(self.G) if self.fused_heads else None, st()))
            if not self.fused_heads:
                # head weight / bias gradients (plain batched GEMM + column sums)
                self.gv["wo"].baddbmm_(H.transpose(1, 2), dlog)
                self.gv["bo"].add_(dlog.sum(dim=1))
            fuse_dx = all_tc and self.dx_fused and self.dx_fusable   # dX = dZ . Wx^T inside the BPTT kernel (second MMA per step)
            if self.bwd_tc:
                gb = (_p(self.st_g[ci]), _p(self.st_c[ci])) if use_store else (None, None)
                _lib.check(lib.tscl_lstm_seq_bwd_tc_dx(self._h, _p(self.Wt), _p(ZG), _p(Cc), _p(dH), _p(self.c_bw),
                                                       _p(dpre), C.c_int32(T), C.c_int64(rc), C.c_int64(R), C.c_int64(r0),
                                                       *gb, _p(dZb), _p(self.Wxt) if fuse_dx else None,
                                                       _p(dXb) if fuse_dx else None, st()))
            else:
                _lib.check(lib.tscl_lstm_seq_bwd(self._h, _p(self.P), _p(ZG), _p(Cc), _p(dH), _p(self.c_bw), _p(dpre),
                                                 C.c_int32(T), C.c_int64(rc), C.c_int64(R), C.c_int64(r0), st()))
            dZ = ZG
            if self.wgrad_tc:
                if use_store:
                    _lib.check(lib.tscl_wgrad_tc(self._h, _p(dZ), _p(dZb), None, _p(self.st_x[ci]), None, _p(self.st_h[ci]),
                                                 _p(self.h_bw), _p(dpre), C.c_int32(T), C.c_int64(rc), C.c_int64(R),
                                                 C.c_int64(r0), _p(self.G), C.c_int32(0), st()))
                else:
                    _lib.check(lib.tscl_wgrad_tc(self._h, _p(dZ), None, _p(X), None, _p(Hp), None, None, None, C.c_int32(T),
                                                 C.c_int64(rc), C.c_int64(R), C.c_int64(r0), _p(self.G), C.c_int32(0), st()))
            else:
                self.gv["wx"].baddbmm_(X.transpose(1, 2), dZ)
                self.gv["wh"].baddbmm_(Hp.transpose(1, 2), dZ)
                self.gv["bl"].add_(dZ.sum(dim=1))
            # dX = dZ . Wx^T: own warp-specialised tcgen05 kernel on the shipping path (tscl_dx_tc); a library GEMM only on
            # the fp32 twin path, for dx > 224 and under TSC_DX_LIBRARY=1 (A/B measurements)
            if all_tc and not fuse_dx and self.dx_own:
                _lib.check(lib.tscl_dx_tc(self._h, _p(dZb), _p(self.Wxt), _p(dXb), C.c_int64(M), st()))
            elif all_tc and not fuse_dx:
                torch.bmm(dZb, self.wx_b.transpose(1, 2), out=dXb)
            elif not all_tc:
                torch.bmm(dZ, self.pv["wx"].transpose(1, 2), out=dX)
            if self.fc_bwd_tc:
                xb = _p(self.st_x[ci]) if use_store else None
                _lib.check(lib.tscl_fc_bwd_tc(self._h, _p(obs0), None if all_tc else _p(X), xb, _p(dX), _p(dXb),
                                              C.c_int64(M), C.c_int64(rc), C.c_int64(R * n_obs), _p(self.G), C.c_int32(0),
                                              st()))
            else:
                _lib.check(lib.tscl_fc_bwd(self._h, _p(obs0), _p(X), _p(dX), C.c_int64(M), C.c_int64(rc),
                                           C.c_int64(R * n_obs), _p(self.G), st()))
            self.kernel_launches += 4 if use_store else 5
        if self.pg is not None:
            _dist.allreduce_sum_(self.G, self.pg)
        _lib.check(lib.tscl_clip_rmsprop(self._h, _p(self.P), _p(self.G), _p(self.MS), _p(self.agent_of),
                                         C.c_float(self.max_grad_norm), C.c_float(lr), C.c_float(self.alpha),
                                         C.c_float(self.eps), _p(self.norms), st()))
        self.kernel_launches += 3
        self.pack_weights()
        # states_bw <- states_fw (agents/policies.py:153); next rollout starts at slot 0
        self.c_bw.copy_(self.c_fw); self.h_bw.copy_(self.h_fw)
        self.obs_hist[0].copy_(self.obs_hist[T])
        self.last_done = bool(self.done_post[-1])
        self.t = 0
        self._acts_ok = [False] * T
