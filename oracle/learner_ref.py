"""CPU ORACLE of the learner math (test infrastructure only): float64 torch/numpy restatement of

  fc / lstm              reference agents/utils.py:66-74, 88-116  (gate order i,f,o,u; done masks c,h)
  FPLstmACPolicy         reference agents/policies.py:191-211     (h = concat(fcw, fcf, fct))
  FcACPolicy             reference agents/policies.py:214-256     (h = relu(fc(concat(fcw, fct))), no state)
  A2C loss               reference agents/policies.py:41-52
  n-step returns         reference agents/utils.py:202-214
  clip + RMSProp         reference agents/policies.py:54-61 (TF1: rms slot starts at 1, eps inside sqrt)

Pinned against the reference where it is pure Python (OnPolicyBuffer, Scheduler: tests/golden/buffers.npz);
the TF1 graph math itself cannot run here (TensorFlow 1.12 is absent) -> parity of those ops is against this
restatement with autograd as the differentiation oracle.
"""
import numpy as np
import torch


def nstep_returns(rs, vs, dones_post, R, gamma):
    """agents/utils.py:202-214."""
    Rs, Advs = [], []
    for r, v, done in zip(rs[::-1], vs[::-1], dones_post[::-1]):
        R = r + gamma * R * (1. - done)
        Rs.append(R)
        Advs.append(R - v)
    return np.array(Rs[::-1]), np.array(Advs[::-1])


def unit_forward(v, lay, u, obs, dones, c, h):
    """obs [T, R, n_obs] -> (pi or value [T, R, ...], H [T, R, h], final c, h).  float64 torch."""
    a = u // 2
    o0 = int(lay.obs_off[a]); nw, nt, nf = int(lay.n_wave[a]), int(lay.n_wait[a]), int(lay.n_fp[a])
    wave = obs[..., o0:o0 + nw]; wait = obs[..., o0 + nw:o0 + nw + nt]; fp = obs[..., o0 + nw + nt:o0 + nw + nt + nf]
    parts = [torch.relu(wave @ v["fcw_w%d" % u] + v["fcw_b%d" % u])]
    if lay.ff > 0:
        parts.append(torch.relu(fp @ v["fcf_w%d" % u] + v["fcf_b%d" % u]))
    if lay.ft > 0:
        parts.append(torch.relu(wait @ v["fct_w%d" % u] + v["fct_b%d" % u]))
    x = torch.cat(parts, -1)
    H = lay.h
    if not getattr(lay, "recurrent", True):      # FcACPolicy: one more fc layer instead of the LSTM
        Hs = torch.relu(x @ v["wx"][u] + v["bl"][u])
        n_out = int(lay.n_a[a]) if u % 2 == 0 else 1
        out = Hs @ v["wo"][u][:, :n_out] + v["bo"][u][:n_out]
        out = torch.softmax(out, -1) if u % 2 == 0 else out.squeeze(-1)
        return out, Hs, c, h
    hs = []
    for t in range(obs.shape[0]):
        keep = 1.0 - dones[t]
        c = c * keep; h = h * keep
        z = x[t] @ v["wx"][u] + h @ v["wh"][u] + v["bl"][u]
        i, f, o, g = (torch.sigmoid(z[:, :H]), torch.sigmoid(z[:, H:2 * H]), torch.sigmoid(z[:, 2 * H:3 * H]),
                      torch.tanh(z[:, 3 * H:]))
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h)
    Hs = torch.stack(hs)
    n_out = int(lay.n_a[a]) if u % 2 == 0 else 1
    out = Hs @ v["wo"][u][:, :n_out] + v["bo"][u][:n_out]
    if u % 2 == 0:
        out = torch.softmax(out, -1)
    else:
        out = out.squeeze(-1)
    return out, Hs, c, h


def a2c_loss(P, lay, obs, acts, Rs, Advs, dones, c0, h0, v_coef, beta):
    """Sum over agents of the per-agent loss averaged over (t, replica).  Returns (loss, per-agent parts)."""
    v = lay.views(P)
    total = 0.0
    parts = []
    for a in range(lay.A):
        pi, _, _, _ = unit_forward(v, lay, 2 * a, obs, dones, c0[2 * a], h0[2 * a])
        val, _, _, _ = unit_forward(v, lay, 2 * a + 1, obs, dones, c0[2 * a + 1], h0[2 * a + 1])
        log_pi = torch.log(torch.clamp(pi, 1e-10, 1.0))
        ent = -(pi * log_pi).sum(-1)
        lp_a = torch.gather(log_pi, -1, acts[..., a].long().unsqueeze(-1)).squeeze(-1)
        pl = -(lp_a * Advs[..., a]).mean()
        vl = ((Rs[..., a] - val) ** 2).mean() * 0.5 * v_coef
        el = -ent.mean() * beta
        total = total + pl + vl + el
        parts.append((pl.item(), vl.item(), el.item()))
    return total, parts


def clip_rmsprop(P, G, MS, agent_of, max_norm, lr, alpha, eps, n_agents):
    P, G, MS = P.copy(), G.copy(), MS.copy()
    norms = np.zeros(n_agents)
    for a in range(n_agents):
        m = agent_of == a
        nrm = np.sqrt((G[m].astype(np.float64) ** 2).sum())
        norms[a] = nrm
        g = G[m] * (max_norm / max(nrm, max_norm) if max_norm > 0 else 1.0)
        MS[m] = alpha * MS[m] + (1 - alpha) * g * g
        P[m] = P[m] - lr * g / np.sqrt(MS[m] + eps)
    return P, MS, norms
