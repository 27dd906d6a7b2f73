"""Golden vectors of the LEARNER produced by executing the reference's own agents/{utils,policies,models}.py
(fc, lstm, LstmACPolicy / FPLstmACPolicy / FcACPolicy / LRQPolicy / DeepQPolicy graph builders, prepare_loss,
IA2C / MA2C forward / add_transition / backward, OnPolicyBuffer) on top of tests/golden/tf1_shim.py — TensorFlow
itself is absent from this container; the shim evaluates the reference-built graphs in float64 (see its header for
the three TF kernels it restates: clip_by_global_norm, RMSProp, Adam).

Outputs (committed): tests/golden/learner_{ma2c,ia2c,fc,iql}.npz
Run:  python tests/golden/gen_learner_golden.py        (needs /root/reference; NOT needed at test time)

Protocol = reference utils.py:142-193,255-291 (Trainer.run / explore): model.reset(), pre-decision done = True on
the first step, n_step x [forward('pv') -> sample -> add_transition], bootstrap forward('v'), backward(R); two
updates, the second batch ends the episode (R = 0).
"""
import configparser
import copy
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import tf1_shim                                                        # noqa: E402

np.bool = bool                      # agents/utils.py:226 uses the alias removed in numpy 1.24
sys.modules["tensorflow"] = tf1_shim.make_module()
import tensorflow as tf                                                # noqa: E402  (the shim)
from agents import models as ref_models, policies as ref_policies      # noqa: E402  (reference)
from agents.utils import OnPolicyBuffer                                # noqa: E402  (reference)

INI = """
[MODEL_CONFIG]
rmsp_alpha = 0.99
rmsp_epsilon = 1e-5
max_grad_norm = 40
gamma = 0.99
lr_init = 5e-4
lr_decay = constant
entropy_coef_init = 0.01
entropy_decay = constant
value_coef = 0.5
num_fw = 128
num_ft = 32
num_lstm = 64
num_fp = 64
batch_size = 6
reward_norm = 3.0
reward_clip = 2.0
epsilon_init = 1.0
epsilon_min = 0.01
epsilon_decay = linear
epsilon_ratio = 0.5
buffer_size = 1000
num_fc = 128
num_h = 64
"""
# two agents with the corner / interior grid shapes, one of them with 4 phases (fixture size: ~160 k weights per agent)
N_WAVE, N_WAIT, N_FP, N_A = [18, 30], [6, 6], [8, 16], [5, 4]
# small max_grad_norm variants exercise the clip branch: rewards are scaled so that agent 2's norm exceeds it


def cfg(**over):
    cp = configparser.ConfigParser()
    cp.read_string(INI)
    for k, v in over.items():
        cp["MODEL_CONFIG"][k] = str(v)
    return cp["MODEL_CONFIG"]


def weights():
    return {v.name: v.numpy() for v in tf.shim_graph().variables.values()}


def put(out, prefix, d):
    for k, v in d.items():
        out["%s/%s" % (prefix, k)] = v


def f32(d):
    return {k: np.asarray(v, np.float32) for k, v in d.items()}


def perturb_biases(rng):
    """non-zero biases so that they matter (the reference initialises them to 0); float32-representable values"""
    for v in tf.shim_graph().variables.values():
        if v.name.endswith("/b"):
            v.assign(rng.normal(0, 0.05, v.example.shape).astype(np.float32))


def run_a2c(kind, max_grad_norm, shapes):
    """kind: 'ma2c' (FPLstmACPolicy) or 'ia2c' (LstmACPolicy), through the reference's model classes.
    Weights and gradients are stored as float32 (the reference's own precision); activations as float64."""
    n_wave, n_wait, n_fp, n_a_ls = shapes
    np.random.seed(7)
    A = len(n_a_ls)
    n_f = n_fp if kind == "ma2c" else [0] * A
    n_s_ls = [w + t + f for w, t, f in zip(n_wave, n_wait, n_f)]
    mc = cfg(max_grad_norm=max_grad_norm)
    if kind == "ma2c":
        model = ref_models.MA2C(n_s_ls, n_a_ls, n_wait, n_f, 1000, mc, seed=0)
    else:
        model = ref_models.IA2C(n_s_ls, n_a_ls, n_wait, 1000, mc, seed=0)
    T = model.n_step
    out = dict(n_wave=np.array(n_wave), n_wait=np.array(n_wait), n_fp=np.array(n_f), n_a=np.array(n_a_ls),
               n_step=T, gamma=0.99, v_coef=0.5, beta=0.01, lr=5e-4, alpha=0.99, eps=1e-5,
               max_grad_norm=float(max_grad_norm), reward_norm=3.0, reward_clip=2.0)
    out["var_names"] = np.array(list(weights().keys()))
    out["var_shapes"] = np.array([str(tuple(v.shape)) for v in weights().values()])
    rng = np.random.default_rng(11)
    perturb_biases(rng)
    put(out, "w0", f32(weights()))
    wts = {i: tf.trainable_variables(scope=model.policy_ls[i].name) for i in range(A)}
    grad_ops = {i: tf.gradients(model.policy_ls[i].loss, wts[i]) for i in range(A)}
    model.reset()
    done = True
    ob = [rng.random(n) * 2 for n in n_s_ls]
    for b in range(2):
        rec = dict(obs=[], acts=[], rew=[], done_pre=[], done_post=[], pi=[], val=[], states=[])
        for t in range(T):
            policy, value = model.forward(ob, done)
            # actions are inputs of the learner: drawn independently of pi so that variants share the trajectory
            action = [int(rng.integers(0, n)) for n in n_a_ls]
            rec["done_pre"].append(float(done))
            reward = rng.normal(0, 4 if b == 0 else 12, A)
            nxt = [rng.random(n) * 2 for n in n_s_ls]
            done = (b == 1 and t == T - 1)
            model.add_transition(ob, action, reward, value, done)
            rec["obs"].append(np.concatenate(ob)); rec["acts"].append(action); rec["rew"].append(reward)
            rec["done_post"].append(float(done))
            pi = np.zeros((A, max(n_a_ls)))
            for i, p in enumerate(policy):
                pi[i, :len(p)] = p
            rec["pi"].append(pi); rec["val"].append(np.array(value, np.float64))
            rec["states"].append(np.stack([p.states_fw for p in model.policy_ls]))
            ob = nxt
        R = [0] * A if done else model.forward(ob, False, 'v')
        rec["next_obs"] = np.concatenate(ob)
        rec["boot"] = np.array(R, np.float64)
        # loss / gradients of every agent on exactly the batch backward() will use (copy of its buffers)
        losses, norms, Rs_all, Adv_all = [], [], [], []
        for i in range(A):
            buf = copy.deepcopy(model.trans_buffer_ls[i])
            obs_b, acts_b, dones_b, Rs, Advs = buf.sample_transition(R[i])
            p = model.policy_ls[i]
            feed = {p.ob_bw: obs_b, p.done_bw: dones_b, p.states: p.states_bw, p.A: acts_b, p.ADV: Advs, p.R: Rs,
                    p.entropy_coef: 0.01}
            vals = model.sess.run([p.loss, p.grad_norm, p.pi, p.v] + grad_ops[i], feed)
            losses.append(float(vals[0])); norms.append(float(vals[1]))
            if b == 0:
                for w, g in zip(wts[i], vals[4:]):
                    out["b0/g/%s" % w.name] = g.astype(np.float32)
                out["b0/pi_bw_%d" % i] = vals[2]; out["b0/v_bw_%d" % i] = vals[3]
            Rs_all.append(Rs); Adv_all.append(Advs)
            out["b%d/dones_bw_%d" % (b, i)] = dones_b.astype(np.float64)
        rec["Rs"] = np.stack(Rs_all, 1); rec["Advs"] = np.stack(Adv_all, 1)
        rec["loss"] = np.array(losses); rec["grad_norm"] = np.array(norms)
        model.backward(R)
        for k, v in rec.items():
            out["b%d/%s" % (b, k)] = np.asarray(v)
        put(out, "w%d" % (b + 1), f32(weights()))
    return out


def run_fc():
    """FcACPolicy (agents/policies.py:214-256) is never instantiated by the reference's models; drive the class
    directly with the same protocol (forward per step, OnPolicyBuffer, backward on the batch)."""
    np.random.seed(9)
    tf.reset_default_graph()
    sess = tf.Session()
    A, T = len(N_A), 6
    pols, bufs = [], []
    for i in range(A):
        p = ref_policies.FcACPolicy(N_WAVE[i], N_A[i], N_WAIT[i], T, n_fc_wave=128, n_fc_wait=32, n_lstm=64,
                                    name='%da' % i)
        p.prepare_loss(0.5, 40.0, 0.99, 1e-5)
        pols.append(p); bufs.append(OnPolicyBuffer(0.99))
    out = dict(n_wave=np.array(N_WAVE), n_wait=np.array(N_WAIT), n_a=np.array(N_A), n_step=T, gamma=0.99, v_coef=0.5,
               beta=0.01, lr=5e-4, alpha=0.99, eps=1e-5, max_grad_norm=40.0)
    out["var_names"] = np.array(list(weights().keys()))
    rng = np.random.default_rng(13)
    perturb_biases(rng)
    put(out, "w0", f32(weights()))
    n_s_ls = [w + t for w, t in zip(N_WAVE, N_WAIT)]
    ob = [rng.random(n) * 2 for n in n_s_ls]
    for b in range(2):
        rec = dict(obs=[], acts=[], rew=[], done_post=[], pi=[], val=[])
        for t in range(T):
            pv = [p.forward(sess, ob[i], False) for i, p in enumerate(pols)]
            action = [int(rng.integers(0, n)) for n in N_A]
            reward = np.clip(rng.normal(0, 1.5, A), -2, 2)
            done = (b == 1 and t == T - 1)
            for i in range(A):
                bufs[i].add_transition(ob[i], action[i], reward[i], pv[i][1], done)
            pi = np.zeros((A, max(N_A)))
            for i, x in enumerate(pv):
                pi[i, :len(x[0])] = x[0]
            rec["obs"].append(np.concatenate(ob)); rec["acts"].append(action); rec["rew"].append(reward)
            rec["done_post"].append(float(done)); rec["pi"].append(pi)
            rec["val"].append(np.array([x[1] for x in pv], np.float64))
            ob = [rng.random(n) * 2 for n in n_s_ls]
        R = [0.0] * A if done else [float(p.forward(sess, ob[i], False, 'v')) for i, p in enumerate(pols)]
        rec["boot"] = np.array(R)
        Rs_all, Adv_all, losses, norms = [], [], [], []
        for i, p in enumerate(pols):
            obs_b, acts_b, dones_b, Rs, Advs = bufs[i].sample_transition(R[i])
            wts = tf.trainable_variables(scope=p.name)
            feed = {p.obs: obs_b, p.A: acts_b, p.ADV: Advs, p.R: Rs, p.entropy_coef: 0.01}
            vals = sess.run([p.loss, p.grad_norm] + tf.gradients(p.loss, wts), feed)
            losses.append(float(vals[0])); norms.append(float(vals[1]))
            for w, g in zip(wts, vals[2:]):
                out["b%d/g/%s" % (b, w.name)] = g.astype(np.float32)
            p.backward(sess, obs_b, acts_b, dones_b, Rs, Advs, 5e-4, 0.01)
            Rs_all.append(Rs); Adv_all.append(Advs)
        rec["Rs"] = np.stack(Rs_all, 1); rec["Advs"] = np.stack(Adv_all, 1)
        rec["loss"] = np.array(losses); rec["grad_norm"] = np.array(norms)
        for k, v in rec.items():
            out["b%d/%s" % (b, k)] = np.asarray(v)
        put(out, "w%d" % (b + 1), f32(weights()))
    return out


def run_iql():
    """LRQPolicy / DeepQPolicy (agents/policies.py:285-389): q-values, TD loss (no target network), clip, Adam —
    three consecutive policy.backward() calls on fixed minibatches; plus IQL.forward / add_transition / backward
    plumbing of agents/models.py:264-376 for the 'lr' model."""
    out = {}
    for kind in ("lr", "dqn"):
        np.random.seed(21); random.seed(21)
        n_s_ls, n_a_ls, n_w_ls = [24, 36], [5, 4], [0, 0]      # n_w = 0: the dqn wait branch has a float layer width
        model = ref_models.IQL(n_s_ls, n_a_ls, n_w_ls, 1000, cfg(batch_size=20, lr_init=1e-4), seed=0,
                               model_type=kind)
        rng = np.random.default_rng(5)
        perturb_biases(rng)
        out["%s/var_names" % kind] = np.array(list(weights().keys()))
        put(out, "%s/w0" % kind, f32(weights()))
        B = model.n_step
        for k in range(3):
            for i, p in enumerate(model.policy_ls):
                obs = rng.random((B, n_s_ls[i])) * 2
                nxt = rng.random((B, n_s_ls[i])) * 2
                acts = rng.integers(0, n_a_ls[i], B)
                rs = rng.normal(0, 1, B)
                dones = rng.random(B) < 0.2
                wts = tf.trainable_variables(scope=p.name)
                feed = {p.S: obs, p.A: acts, p.S1: nxt, p.DONE: dones, p.R: rs}
                vals = model.sess.run([p.loss, p.grad_norm, p.qvalues] + tf.gradients(p.loss, wts), feed)
                pre = "%s/k%d/a%d" % (kind, k, i)
                out[pre + "/obs"], out[pre + "/next_obs"], out[pre + "/acts"] = obs, nxt, acts
                out[pre + "/rs"], out[pre + "/dones"] = rs, dones
                out[pre + "/loss"], out[pre + "/grad_norm"], out[pre + "/q"] = vals[0], vals[1], vals[2]
                for w, g in zip(wts, vals[3:]):
                    out[pre + "/g/" + w.name] = g.astype(np.float32)
                p.backward(model.sess, obs, acts, nxt, dones, rs, 1e-4)
            put(out, "%s/w%d" % (kind, k + 1), f32(weights()))
        if kind == "lr":       # model-level plumbing: epsilon schedule + greedy action = argmax q
            ob = [rng.random(n) for n in n_s_ls]
            acts, qs = model.forward(ob, mode='act')
            out["lr/plumb_obs"] = np.concatenate(ob)
            out["lr/plumb_q"] = np.concatenate([np.atleast_1d(q) for q in qs])
            out["lr/plumb_act"] = np.array(acts)
            eps = []
            for _ in range(5):
                model.forward(ob, mode='explore')
                eps.append(model.eps_scheduler.val * (1 - model.eps_scheduler.n / model.eps_scheduler.N))
            out["lr/eps"] = np.array(eps)
    return out


def main():
    ma = (N_WAVE, N_WAIT, N_FP, N_A)
    o = run_a2c("ma2c", 40, ma)
    del_keys = [k for k in o if k.startswith("w1/")]      # w2 (after two updates) also pins the carried optimizer state
    for k in del_keys:
        del o[k]
    c = run_a2c("ma2c", 0.05, ma)                          # clip branch active: same trajectory, first update only
    for k in c:
        if k.startswith("w1/"):
            o["clip/" + k] = c[k]
    o["clip/max_grad_norm"] = 0.05
    o["clip/b0/grad_norm"] = c["b0/grad_norm"]
    assert np.array_equal(c["b0/obs"], o["b0/obs"]) and np.array_equal(c["b0/Advs"], o["b0/Advs"])
    np.savez_compressed(os.path.join(HERE, "learner_ma2c.npz"), **o)
    print("ma2c loss", o["b0/loss"], "grad_norm", o["b0/grad_norm"], o["b1/grad_norm"])
    o = run_a2c("ia2c", 40, ([24], [6], [0], [5]))
    for k in [k for k in o if k.startswith("w1/")]:
        del o[k]
    np.savez_compressed(os.path.join(HERE, "learner_ia2c.npz"), **o)
    print("ia2c loss", o["b0/loss"], "grad_norm", o["b0/grad_norm"], o["b1/grad_norm"])
    o = run_fc()
    np.savez_compressed(os.path.join(HERE, "learner_fc.npz"), **o)
    print("fc loss", o["b0/loss"], "grad_norm", o["b0/grad_norm"])
    o = run_iql()
    np.savez_compressed(os.path.join(HERE, "learner_iql.npz"), **o)
    print("iql lr loss", o["lr/k0/a0/loss"], "dqn loss", o["dqn/k0/a0/loss"])


if __name__ == "__main__":
    main()
