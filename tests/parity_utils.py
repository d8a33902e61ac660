"""GPU parity harness: drives the CUDA drop-in and the CPU oracle on identical seeded inputs and
identical injected sampling randomness, step by step, and returns error metrics.
Used by the -m gpu tests and by __graft_entry__.smoke()."""
from __future__ import annotations

import ctypes as C
import tempfile

import numpy as np
import torch

from oracle.dqn_oracle import OracleLearner, numpy_adam, synthetic_transitions
from tests.recipes import max_rel, perturb_target


def tensor_sizes(net):
    return [(k, int(v.numel())) for k, v in net.state_dict().items()]


def per_tensor_max_rel(flat_a, flat_b, sizes):
    out, off = {}, 0
    for k, n in sizes:
        out[k] = max_rel(flat_a[off:off + n], flat_b[off:off + n])
        off += n
    return out


def flat_sd(net):
    return np.concatenate([v.detach().cpu().numpy().ravel() for v in net.state_dict().values()])


def make_pair(algo, D, B, cap, fill, seed, soft=True, target_freq=30000, tmpdir=None, activation="relu", body="macro", gpu="0"):
    """(oracle learner, CUDA agent) with identical weights and identical replay contents."""
    from multimodal_drl_rmc_b200 import macro_config
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    orc = OracleLearner(algo, D, 8, B, cap, soft=soft, target_freq=target_freq, activation=activation, body=body)
    perturb_target(orc.target, seed + 100)
    tmp = tmpdir or tempfile.mkdtemp(prefix="rmc_parity_")
    agent = macro_config.make_agent(algo, D, B, cap, save_dir=tmp + "/", log_dir=tmp + "/",
                                    target_soft_update=soft, target_update_freq=target_freq,
                                    activation="hybrid" if body == "hybrid" else activation, gpu=gpu)
    agent.online_network.load_state_dict({k: v.clone() for k, v in orc.online.state_dict().items()})
    agent.target_network.load_state_dict({k: v.clone() for k, v in orc.target.state_dict().items()})
    obs, act, rew, done, nxt = synthetic_transitions(fill, D, 20251018 + seed)
    for i in range(fill):
        orc.store([obs[i]], [int(act[i])], [float(rew[i])], [bool(done[i])], [nxt[i]])
    # the drop-in receives the same stream, in uneven chunks (exercises ring wrap + multi-row pushes)
    i = 0
    chunk = 1
    while i < fill:
        j = min(fill, i + chunk)
        agent.store_transitions(obs[i:j], act[i:j].tolist(), rew[i:j].tolist(), done[i:j].astype(bool).tolist(), nxt[i:j], None)
        i, chunk = j, (chunk * 3) % 97 + 1
    return orc, agent


class LazyRows:
    """``SumTree.data`` / deque stand-in for bulk-filled oracles: the transition tuple of slot i is built on demand from
    the arrays (a materialised object array of a million tuples costs ~0.5 GB and seconds to build)."""

    def __init__(self, obs, act, rew, done, nxt):
        self.obs, self.act, self.rew, self.done, self.nxt = obs, act, rew, done, nxt

    def __len__(self):
        return len(self.act)

    def __getitem__(self, i):
        i = int(i)
        return (self.obs[i], int(self.act[i]), float(self.rew[i]), bool(self.done[i] != 0), self.nxt[i])


def make_pair_bulk(algo, D, B, cap, seed, pri_seed=7, activation="relu"):
    """(oracle learner, CUDA agent) with identical weights and a replay FILLED TO CAPACITY (size = cap, data_pointer = 0)
    the vectorised way bench.py fills it: the oracle's own data structures are written directly (leaves <- seeded float32
    priorities, inner nodes rebuilt level by level -- exact sums --, arg-max / arg-min leaf indices), the drop-in gets the
    rows through its bulk push and the same priorities through rmc_replay_set_priorities.  For the sizes the headline
    numbers are quoted on (cap = size = 1,000,000), where a transition-by-transition fill of the python oracle takes minutes."""
    from multimodal_drl_rmc_b200 import _lib, macro_config
    from multimodal_drl_rmc_b200.synthetic import seeded_priorities
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    orc = OracleLearner(algo, D, 8, B, cap, activation=activation)
    perturb_target(orc.target, seed + 100)
    tmp = tempfile.mkdtemp(prefix="rmc_parity_")
    agent = macro_config.make_agent(algo, D, B, cap, save_dir=tmp + "/", log_dir=tmp + "/", activation=activation)
    agent.online_network.load_state_dict({k: v.clone() for k, v in orc.online.state_dict().items()})
    agent.target_network.load_state_dict({k: v.clone() for k, v in orc.target.state_dict().items()})
    obs, act, rew, done, nxt = synthetic_transitions(cap, D, 20251018 + seed)
    rows = LazyRows(obs, act, rew, done, nxt)
    ring = agent.replay_memory_buffer._ring
    ring.push_host(obs, act, rew, done, nxt)
    if orc.per:
        pri = seeded_priorities(cap, pri_seed + seed)
        t = orc.replay.tree
        t.data = rows
        t.size, t.data_pointer = cap, 0
        t.tree[cap - 1:] = pri
        level = int(np.floor(np.log2(cap - 1))) if cap > 1 else 0
        for L in range(level, -1, -1):
            first, last = (1 << L) - 1, min((1 << (L + 1)) - 2, cap - 2)
            if first <= last:
                idx = np.arange(first, last + 1)
                t.tree[idx] = t.tree[2 * idx + 1] + t.tree[2 * idx + 2]
        leaves = t.tree[cap - 1:]
        t.arg_max, t.arg_min = int(np.argmax(leaves)) + cap - 1, int(np.argmin(leaves)) + cap - 1
        pt = torch.as_tensor(pri, device=agent.device)
        _lib.check(_lib.lib().rmc_replay_set_priorities(ring.handle, pt.data_ptr(), cap, _lib.stream_ptr(agent.device.index)))
        torch.cuda.synchronize()
    else:
        orc.replay.buf = rows            # deque stand-in: indexable by age (oldest first == slot order for a ring that has just filled)
    return orc, agent


def gpu_out(agent, name, dtype=torch.float32):
    return agent._lh.output(name, dtype).cpu().numpy()


def adam_state(orc):
    m = np.concatenate([orc.opt.state[p]["exp_avg"].numpy().ravel() if p in orc.opt.state else np.zeros(p.numel(), np.float32) for p in orc.online.parameters()])
    v = np.concatenate([orc.opt.state[p]["exp_avg_sq"].numpy().ravel() if p in orc.opt.state else np.zeros(p.numel(), np.float32) for p in orc.online.parameters()])
    return m, v


def ulp_err(a, b, floor):
    """|a - b| in units of the float32 spacing at max(|b|, floor), max over all elements."""
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    sp = np.spacing(np.maximum(np.abs(np.asarray(b, np.float32)), np.float32(floor))).astype(np.float64)
    return float(np.max(np.abs(a64 - b64) / sp))


def f64_gradients(orc, online64, target64, tr):
    """Gradients of the oracle's own minibatch (trace ``tr``) recomputed in float64 from float64 copies of the PRE-step
    weights: the exact-arithmetic answer both fp32 implementations approximate (dqn/agent.py:245-272, same formulas)."""
    o, n = torch.as_tensor(tr["obs"]).double(), torch.as_tensor(tr["nxt"]).double()
    a = torch.as_tensor(tr["act"])
    r, d = torch.as_tensor(tr["rew"]).double(), torch.as_tensor(tr["done"]).double()
    with torch.no_grad():
        q_next_tgt = target64(n)
        if orc.double:
            q_sel = torch.gather(q_next_tgt, 1, online64(n).argmax(dim=1, keepdim=True))
        else:
            q_sel = q_next_tgt.max(dim=1, keepdim=True)[0]
        y = r + (1 - d) * float(np.float32(orc.gamma)) * q_sel
    q_sa = torch.gather(online64(o), 1, a)
    hub = torch.nn.functional.smooth_l1_loss(q_sa, y, reduction="none")
    if orc.per:
        w = torch.as_tensor(tr["is_w"].astype(np.float32)).double().unsqueeze(-1)
        loss = torch.mean(w * hub)
    else:
        loss = torch.mean(hub)
    online64.zero_grad()
    loss.backward()
    return np.concatenate([p.grad.numpy().ravel() for _, p in online64.named_parameters()])


def run_parity_case(algo, D, B, cap, fill, steps, seed, soft=True, target_freq=30000, resync_tree=True, activation="relu", body="macro", pair=None,
                    f64_truth=False):
    """Steps the oracle and the CUDA drop-in side by side.  Weights are checked three ways so that EVERY element is covered:
      * vs the oracle at 1e-5 on the elements whose gradient is well conditioned for Adam (|g| >= 1e-6; the fraction is
        returned as ``well_conditioned_frac``), and bounded by lr per step on the rest;
      * Adam closure on ALL elements: the device's post-step (p, m, v) against numpy_adam (the bit-matched restatement of
        torch.optim.Adam) applied to the DEVICE's own pre-step (p, m, v) and gradients -- in ulps -- and the Polyak / hard
        target update against its formula bit for bit;
      * Adam moments m, v vs the oracle on ALL elements (they are well conditioned everywhere), and elements whose oracle
        gradient was exactly 0 in every step so far (dead units) must equal the oracle's weights bit for bit."""
    from multimodal_drl_rmc_b200 import _lib
    orc, agent = pair if pair is not None else make_pair(algo, D, B, cap, fill, seed, soft, target_freq, activation=activation, body=body)
    per = orc.per
    sizes = tensor_sizes(orc.online)
    rng = np.random.default_rng(seed + 1)
    res = dict(nodes_equal=True, tree_equal=True, max_rel_q=0.0, max_rel_loss=0.0, max_rel_isw=0.0,
               max_rel_grads=0.0, max_rel_weights=0.0, max_rel_target=0.0, max_pri_ulp=0.0, worst_grad="", worst_w="")
    if per:
        res["tree_equal"] = bool(np.array_equal(agent.replay_memory_buffer.replay_buffer.tree, orc.replay.tree.tree))
    well = None
    always_zero = None
    res.update(adam_closure_ulp=0.0, adam_m_ulp=0.0, adam_v_ulp=0.0, polyak_bitexact=True, max_rel_m=0.0, max_rel_v=0.0,
               zero_grad_exact=True, zero_grad_weights_bitexact=True)
    k_pol = np.float32(agent.target_soft_update_tau * agent.n_env)
    k_1m = np.float32(1.0 - agent.target_soft_update_tau * agent.n_env)
    for s in range(steps):
        step_no = 1000 * s + 17
        orc.step = agent.step = step_no
        tr = {}
        lh = agent._lh
        pre = {k: lh.get_params(kind).cpu().numpy() for k, kind in (("p", _lib.ONLINE), ("t", _lib.TARGET), ("m", _lib.ADAM_M), ("v", _lib.ADAM_V))}
        if f64_truth:
            import copy
            online64, target64 = copy.deepcopy(orc.online).double(), copy.deepcopy(orc.target).double()
        if per:
            u = rng.random(B)
            orc.learn(u=u, trace=tr)
            agent.learn(u=u)
        else:
            idx = rng.permutation(len(orc.replay.buf))[:B].astype(np.int64)
            orc.learn(indices=[int(i) for i in idx], trace=tr)
            agent.learn(indices=idx)
        # ---- per-sample products
        if per:
            nodes = gpu_out(agent, "nodes", torch.int64)
            res["nodes_equal"] &= bool(np.array_equal(nodes, tr["nodes"]))
            res["max_rel_isw"] = max(res["max_rel_isw"], max_rel(gpu_out(agent, "is_w"), tr["is_w"].astype(np.float32)))
        res["max_rel_q"] = max(res["max_rel_q"], max_rel(gpu_out(agent, "q_sa"), tr["q_sa"].reshape(-1)),
                               max_rel(gpu_out(agent, "y"), tr["y"].reshape(-1)))
        A = 8
        qn = gpu_out(agent, "q").reshape(B, -1)[:, :A]
        res["max_rel_q"] = max(res["max_rel_q"], max_rel(qn, tr["q"]))
        res["max_rel_loss"] = max(res["max_rel_loss"], abs(agent.last_loss() - tr["loss"]) / max(abs(tr["loss"]), 1e-30))
        # ---- gradients (torch state_dict order)
        g_gpu = agent._lh.get_params(_lib.GRADS).cpu().numpy()
        g_ref = np.concatenate([tr["grads"][k].ravel() for k, _ in orc.online.named_parameters()])
        pt = per_tensor_max_rel(g_gpu, g_ref, sizes)
        if "fc_val.bias" in pt:
            # a ONE-element tensor: d(fc_val.bias) = sum_i g_i, a cancelling sum of the per-sample loss coefficients -- its own
            # magnitude says nothing about the conditioning of the sum, so it is measured against sum_i |g_i|
            off = sum(n for k, n in sizes[:[k for k, _ in sizes].index("fc_val.bias")])
            scale = float(np.sum(np.abs(gpu_out(agent, "gcoef").astype(np.float64))))
            pt["fc_val.bias"] = abs(float(g_gpu[off]) - float(g_ref[off])) / max(scale, 1e-30)
        worst = max(pt, key=pt.get)
        if pt[worst] > res["max_rel_grads"]:
            res["max_rel_grads"], res["worst_grad"] = pt[worst], worst
        if f64_truth:      # both fp32 results against exact arithmetic: whose rounding is the difference above?
            g64 = f64_gradients(orc, online64, target64, tr)
            e_gpu, e_ref = per_tensor_max_rel(g_gpu, g64, sizes), per_tensor_max_rel(g_ref, g64, sizes)
            e_gpu.pop("fc_val.bias", None), e_ref.pop("fc_val.bias", None)
            res["gpu_grads_vs_f64"] = max(res.get("gpu_grads_vs_f64", 0.0), max(e_gpu.values()))
            res["ref_grads_vs_f64"] = max(res.get("ref_grads_vs_f64", 0.0), max(e_ref.values()))
            res["grads_vs_f64_per_tensor"] = {k: (float("%.3g" % e_gpu[k]), float("%.3g" % e_ref[k])) for k in e_gpu}
        # ---- priorities / tree
        if per:
            p_ref = np.power(np.minimum(tr["abs_td"].reshape(-1) + np.float32(1e-4), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
            p_gpu = gpu_out(agent, "pri")
            # the |td| -> p map is checked on the device's own |td| (|td| itself is a 1e-5-tolerance item)
            td_gpu = gpu_out(agent, "abs_td")
            res["max_rel_q"] = max(res["max_rel_q"], max_rel(td_gpu, tr["abs_td"].reshape(-1)))
            p_same = np.power(np.minimum(td_gpu + np.float32(1e-4), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
            ulp = np.abs(p_gpu.astype(np.float64) - p_same.astype(np.float64)) / np.spacing(p_same).astype(np.float64)
            res["max_pri_ulp"] = max(res["max_pri_ulp"], float(ulp.max()))
            if resync_tree:   # make the trees bit-identical again (1-ulp pow / |td| differences), then compare
                dev = agent.device
                n_t = torch.as_tensor(tr["nodes"], device=dev)
                p_t = torch.as_tensor(p_ref, device=dev)
                _lib.check(_lib.lib().rmc_per_update(agent.replay_memory_buffer._ring.handle, n_t.data_ptr(), p_t.data_ptr(),
                                                     B, _lib.stream_ptr()))
                t_gpu = agent.replay_memory_buffer.replay_buffer.tree
                res["tree_equal"] &= bool(np.array_equal(t_gpu, orc.replay.tree.tree))
                st = agent.replay_memory_buffer._ring.stats()
                res["tree_equal"] &= (st.total_priority == orc.replay.tree.total and st.max_priority == orc.replay.tree.max_leaf
                                      and st.min_priority == orc.replay.tree.min_leaf)
        # ---- Adam closure on the device's own inputs, ALL elements (independent of how well conditioned the gradient is)
        t_adam = agent._adam_t
        p_exp, m_exp, v_exp = numpy_adam(pre["p"], g_gpu, pre["m"], pre["v"], t_adam, lr=agent.lr)
        post_p, post_m, post_v = (lh.get_params(kind).cpu().numpy() for kind in (_lib.ONLINE, _lib.ADAM_M, _lib.ADAM_V))
        res["adam_closure_ulp"] = max(res["adam_closure_ulp"], ulp_err(post_p, p_exp, 1e-4))
        res["adam_m_ulp"] = max(res["adam_m_ulp"], ulp_err(post_m, m_exp, 1e-30))
        res["adam_v_ulp"] = max(res["adam_v_ulp"], ulp_err(post_v, v_exp, 1e-30))
        m_ref, v_ref = adam_state(orc)
        res["max_rel_m"] = max(res["max_rel_m"], max(per_tensor_max_rel(post_m, m_ref, sizes).values()))
        res["max_rel_v"] = max(res["max_rel_v"], max(per_tensor_max_rel(post_v, v_ref, sizes).values()))
        # exact zeros of the oracle's gradient (dead units: every ReLU mask of the unit is off) must be exact zeros here
        zero = g_ref == 0
        res["zero_grad_exact"] &= bool(np.all(g_gpu[zero] == 0))
        always_zero = zero if always_zero is None else (always_zero & zero)
        # ---- target sync, then weights
        orc.sync_target()
        agent.update_target_network()
        post_t = lh.get_params(_lib.TARGET).cpu().numpy()
        hard = (not agent.target_soft_update) and (agent.step % (agent.update_target_frequency // agent.n_env) == 0)
        t_exp = post_p if hard else ((k_pol * post_p + k_1m * pre["t"]).astype(np.float32) if agent.target_soft_update else pre["t"])
        res["polyak_bitexact"] &= bool(np.array_equal(post_t, t_exp))
        # Post-Adam weights.  Adam divides by (|g| + eps): an element whose gradient is below ~1e-6 is
        # ill-conditioned (d update / d g = lr*eps/(|g|+eps)^2 up to 1e4), so two correct fp32 summation
        # orders legitimately differ there by up to ~lr (the reference differs from a numpy
        # restatement of itself the same way).  1e-5 is asserted on the well-conditioned elements and the
        # rest is bounded by lr per step.
        w_gpu, w_ref = flat_sd(agent.online_network), flat_sd(orc.online)
        well = well & (np.abs(g_ref) >= 1e-6) if s else (np.abs(g_ref) >= 1e-6)
        w_pt = per_tensor_max_rel(np.where(well, w_gpu, w_ref), w_ref, sizes)
        worst = max(w_pt, key=w_pt.get)
        if w_pt[worst] > res["max_rel_weights"]:
            res["max_rel_weights"], res["worst_w"] = w_pt[worst], worst
        res["max_abs_weights_all"] = max(res.get("max_abs_weights_all", 0.0), float(np.max(np.abs(w_gpu - w_ref))))
        res["well_conditioned_frac"] = float(well.mean())
        res["zero_grad_frac"] = float(always_zero.mean())
        res["zero_grad_weights_bitexact"] &= bool(np.array_equal(w_gpu[always_zero], w_ref[always_zero]))
        t_pt = per_tensor_max_rel(flat_sd(agent.target_network), flat_sd(orc.target), sizes)
        res["max_rel_target"] = max(res["max_rel_target"], max(t_pt.values()))
    return res
