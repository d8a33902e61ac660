"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DQN learner hot path.

This file is a CPU restatement (torch-CPU + numpy + python loops, i.e. the same
numerics stack the reference executes on) of the replay-minibatch update of
youcefMehamlia/Multimodal-DRL-RMC.  It exists so that the CUDA path can be checked on
a box where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; the product package never does (and fails loudly without its
CUDA library instead of falling back to this).

Parity pin: the reference holds no tests or golden vectors of its own for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the *reference itself*,
generated in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/dqn`` unmodified through ``oracle/refharness.py``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` asserts this file reproduces them
bit-for-bit (same torch/numpy build) -- parity pinned on reference-generated fixtures.

Every function cites the reference lines it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import random as _pyrandom
from collections import deque

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------
# Sum tree  (reference: dqn/utils/sum_tree.py)
# --------------------------------------------------------------------------------------
class OracleSumTree:
    """Array-heap binary sum tree over ``capacity`` leaves.

    Layout (dqn/utils/sum_tree.py:6-13): float64 array of ``2*capacity-1`` nodes, root at
    0, children of ``n`` at ``2n+1`` / ``2n+2``, leaves occupy ``[capacity-1, 2*capacity-2]``;
    ring cursor ``data_pointer``; ``size`` saturates at capacity; the arg-max / arg-min leaf
    are tracked by index and both start at leaf 0.
    """

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.tree = np.zeros(2 * self.capacity - 1)
        self.data = np.zeros(self.capacity, dtype=object)
        self.data_pointer = 0
        self.size = 0
        self.arg_max = self.capacity - 1
        self.arg_min = self.capacity - 1

    # dqn/utils/sum_tree.py:15-32
    def assign(self, node: int, priority) -> None:
        first_leaf = self.capacity - 1
        prev_max = self.tree[self.arg_max]
        prev_min = self.tree[self.arg_min]
        delta = priority - self.tree[node]
        self.tree[node] = priority
        # running extreme bookkeeping; an overwritten extreme triggers a rescan of the
        # filled leaves (sum_tree.py:21-28)
        if priority >= prev_max:
            self.arg_max = node
        elif node == self.arg_max:
            self.arg_max = int(np.argmax(self.tree[first_leaf:first_leaf + self.size])) + first_leaf
        if priority <= prev_min:
            self.arg_min = node
        elif node == self.arg_min:
            self.arg_min = int(np.argmin(self.tree[first_leaf:first_leaf + self.size])) + first_leaf
        # push the difference up to the root (sum_tree.py:30-32)
        while node != 0:
            node = (node - 1) // 2
            self.tree[node] += delta

    # dqn/utils/sum_tree.py:34-40
    def push(self, priority, item) -> None:
        node = self.data_pointer + self.capacity - 1
        self.data[self.data_pointer] = item
        self.data_pointer = (self.data_pointer + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)  # size is bumped BEFORE the update
        self.assign(node, priority)

    # dqn/utils/sum_tree.py:42-61
    def descend(self, v):
        node = 0
        n_nodes = len(self.tree)
        while True:
            left = 2 * node + 1
            if left >= n_nodes:
                break
            if v <= self.tree[left]:
                node = left
            else:
                v -= self.tree[left]
                node = left + 1
        return node, self.tree[node], self.data[node - self.capacity + 1]

    # dqn/utils/sum_tree.py:63-73
    @property
    def total(self):
        return self.tree[0]

    @property
    def max_leaf(self):
        return self.tree[self.arg_max]

    @property
    def min_leaf(self):
        return self.tree[self.arg_min]


# --------------------------------------------------------------------------------------
# Replay memories  (reference: dqn/replay_memory.py)
# --------------------------------------------------------------------------------------
class OracleUniformReplay:
    """dqn/replay_memory.py:24-39 -- bounded deque, sampling without replacement."""

    def __init__(self, capacity: int, batch: int):
        self.capacity, self.batch = int(capacity), int(batch)
        self.buf = deque(maxlen=self.capacity)

    def store(self, obses, actions, rews, dones, new_obses):
        """Generator, exactly like the reference: must be iterated; yields env index on done."""
        for e, row in enumerate(zip(obses, actions, rews, dones, new_obses)):
            self.buf.append(tuple(row))
            if row[3]:
                yield e

    def sample(self, step=None, indices=None):
        if indices is not None:  # injected indices (test hook; same rows random.sample would pick)
            return [self.buf[i] for i in indices]
        return _pyrandom.sample(self.buf, self.batch)


class OraclePrioritizedReplay:
    """dqn/replay_memory.py:43-98 -- proportional PER with stratified sum-tree sampling."""

    def __init__(self, capacity: int, batch: int, beta_steps):
        self.capacity, self.batch = int(capacity), int(batch)
        self.tree = OracleSumTree(self.capacity)
        # constants: dqn/replay_memory.py:49-54
        self.eps = 0.0001
        self.alpha = 0.6
        self.beta0, self.beta1 = 0.4, 1.0
        self.beta_steps = beta_steps
        self.p_cap = 1.0

    # dqn/replay_memory.py:56-67
    def store(self, obses, actions, rews, dones, new_obses):
        p = self.tree.max_leaf  # read once per call
        if p == 0:
            p = self.p_cap
        for e, row in enumerate(zip(obses, actions, rews, dones, new_obses)):
            self.tree.push(p, tuple(row))
            if row[3]:
                yield e

    def beta(self, step):
        return np.interp(step, [0, self.beta_steps], [self.beta0, self.beta1])

    # dqn/replay_memory.py:69-92
    def sample(self, step, u=None):
        """``u``: optional injected uniforms in [0,1) (float64, one per sample).  With ``u`` the
        draw is ``lo + (hi-lo)*u`` which is bitwise what ``np.random.uniform(lo, hi)`` computes
        from the same ``random_sample()`` value (SURVEY.md finding 7)."""
        t = self.tree
        seg = t.total / self.batch
        beta = self.beta(step)
        max_w = pow(t.size * (t.min_leaf / t.total), -beta)
        weights, nodes, rows = [], [], []
        for i in range(self.batch):
            lo, hi = seg * i, seg * (i + 1)
            v = np.random.uniform(lo, hi) if u is None else lo + (hi - lo) * u[i]
            node, p, row = t.descend(v)
            weights.append(pow(t.size * (p / t.total), -beta) / max_w)
            nodes.append(node)
            rows.append(row)
        return weights, nodes, rows

    # dqn/replay_memory.py:94-98
    def write_back(self, nodes, abs_td):
        pri = list(np.power(np.minimum(abs_td + self.eps, self.p_cap), self.alpha))
        for node, p in zip(nodes, pri):
            self.tree.assign(node, p)

    def write_back_priorities(self, nodes, pri):
        """Tree update from already-formed float32 priorities (the bit-exact half of the
        contract; the |td| -> p map is a 1-ulp tolerance item, SURVEY.md finding 8)."""
        for node, p in zip(nodes, pri):
            self.tree.assign(node, p)


# --------------------------------------------------------------------------------------
# Networks  (reference: dqn/network.py + the macro network_config)
# --------------------------------------------------------------------------------------
def macro_body(obs_dim: int, hidden=(256, 128), activation="relu"):
    """env/custom_env/macro with lane/dqn_config.py:58-104 (Linear-ReLU-Linear-ReLU); activation="elu" is the same body
    with the repo-HEAD activation (``ACTIVATION = nn.ELU()``, env/dqn_config.py:175)."""
    act = nn.ELU if activation == "elu" else nn.ReLU
    return nn.Sequential(nn.Linear(obs_dim, hidden[0]), act(),
                         nn.Linear(hidden[0], hidden[1]), act())


class OracleHybridBody(nn.Module):
    """Restatement of the repo-HEAD body (env/dqn_config.py:66-143 ``TwoStreamHybridNetwork`` as configured by
    ``network_config`` :148-193): state = [macro(14) | grid(2,27,5) flattened]; grid -> Conv2d(2,32,3,s(1,1),p1) -> act ->
    Conv2d(32,64,3,s(2,1),p1) -> act -> Conv2d(64,64,3,s(2,2),p1) -> act -> flatten; cat([flatten, macro]) ->
    Linear(1358,512) -> act -> Linear(512,256) -> act.  Attribute names follow the reference so that the state_dict keys
    equal the ``.pack`` checkpoint keys (``net.cnn_stream.0.weight`` ... ``net.dense_stream.2.bias``)."""

    def __init__(self, macro_len=14, micro_shape=(2, 27, 5), cnn=((32, (1, 1)), (64, (2, 1)), (64, (2, 2))), dense=(512, 256),
                 activation="elu"):
        super().__init__()
        act = nn.ELU() if activation == "elu" else nn.ReLU()       # ONE shared module instance, as in the reference
        self.macro_len, self.micro_shape = int(macro_len), tuple(micro_shape)
        layers, ch = [], self.micro_shape[0]
        for out_ch, stride in cnn:                                  # dqn_config.py:84-93
            layers += [nn.Conv2d(ch, out_ch, kernel_size=(3, 3), stride=stride, padding=(1, 1)), act]
            ch = out_ch
        self.cnn_stream = nn.Sequential(*layers)
        with torch.no_grad():                                       # dqn_config.py:98-105
            flat = self.cnn_stream(torch.zeros(1, *self.micro_shape)).flatten(start_dim=1).shape[1]
        layers, width = [], flat + self.macro_len
        for out_f in dense:                                         # dqn_config.py:108-114
            layers += [nn.Linear(width, out_f), act]
            width = out_f
        self.dense_stream = nn.Sequential(*layers)
        self.fc_out_dim = width

    def forward(self, x):                                           # dqn_config.py:119-143
        macro, grid = x[:, :self.macro_len], x[:, self.macro_len:].view(-1, *self.micro_shape)
        feat = self.cnn_stream(grid).flatten(start_dim=1)
        return self.dense_stream(torch.cat([feat, macro], dim=1))


class OracleQNet(nn.Module):
    """dqn/network.py:50-74 (plain head ``fc_out``) and :77-117 (dueling ``fc_val``/``fc_adv``).

    Module attribute names follow the reference so ``state_dict()`` keys equal the ``.pack``
    checkpoint keys (``net.0.weight`` ... ``fc_adv.bias``).  Construction order (body, then
    val, then adv / out) matches the reference so that a seeded default init is identical."""

    def __init__(self, obs_dim: int, n_actions: int, dueling: bool, hidden=(256, 128), activation="relu", body="macro"):
        super().__init__()
        self.dueling = bool(dueling)
        if body == "hybrid":
            self.net = OracleHybridBody(activation=activation)
            hidden = (None, self.net.fc_out_dim)
        else:
            self.net = macro_body(obs_dim, hidden, activation)
        if self.dueling:
            self.fc_val = nn.Linear(hidden[1], 1)
            self.fc_adv = nn.Linear(hidden[1], n_actions)
        else:
            self.fc_out = nn.Linear(hidden[1], n_actions)

    def forward(self, s):
        z = self.net(s)
        if not self.dueling:
            return self.fc_out(z)
        val, adv = self.fc_val(z), self.fc_adv(z)
        return torch.add(val, (adv - adv.mean(dim=1, keepdim=True)))  # network.py:83

    def greedy(self, obses):
        """network.py:67-74 (argmax of Q) / :110-117 (dueling: argmax of RAW advantages)."""
        x = torch.as_tensor(obses, dtype=torch.float32)
        z = self.net(x)
        scores = self.fc_adv(z) if self.dueling else self.fc_out(z)
        return torch.argmax(scores, dim=1).detach().tolist()


# --------------------------------------------------------------------------------------
# Learner  (reference: dqn/agent.py)
# --------------------------------------------------------------------------------------
class OracleLearner:
    """One agent's learner state + the three ``learn()`` variants of dqn/agent.py.

    algo (dqn/agent.py:275-320):
      "DQNAgent"                  uniform replay, plain head, max-target     (:166-185)
      "DoubleDQNAgent"            uniform replay, plain head, double-DQN     (:204-226)
      "DuelingDoubleDQNAgent"     uniform replay, dueling head, double-DQN
      "PerDuelingDoubleDQNAgent"  PER, dueling head, double-DQN, IS weights  (:245-272)
    """

    ALGOS = {
        "DQNAgent": dict(per=False, dueling=False, double=False),
        "DoubleDQNAgent": dict(per=False, dueling=False, double=True),
        "DuelingDoubleDQNAgent": dict(per=False, dueling=True, double=True),
        "PerDuelingDoubleDQNAgent": dict(per=True, dueling=True, double=True),
    }

    def __init__(self, algo: str, obs_dim: int, n_actions: int, batch: int, capacity: int, *,
                 lr=1e-4, gamma=0.99, tau=1e-3, soft=True, target_freq=30000, eps_decay=2e6,
                 n_env=1, activation="relu", body="macro"):
        f = self.ALGOS[algo]
        self.algo, self.per, self.dueling, self.double = algo, f["per"], f["dueling"], f["double"]
        self.obs_dim, self.n_actions, self.batch, self.capacity = obs_dim, n_actions, batch, capacity
        self.lr, self.gamma, self.tau, self.soft = lr, gamma, tau, soft
        self.target_freq, self.n_env, self.step = target_freq, n_env, 0
        self.replay = (OraclePrioritizedReplay(capacity, batch, eps_decay) if self.per
                       else OracleUniformReplay(capacity, batch))
        # construction order online -> target, as in agent.py:282-283 etc.
        self.online = OracleQNet(obs_dim, n_actions, self.dueling, activation=activation, body=body)
        self.target = OracleQNet(obs_dim, n_actions, self.dueling, activation=activation, body=body)
        self.opt = torch.optim.Adam(self.online.parameters(), lr=lr)   # network.py:17,56
        self.huber = nn.SmoothL1Loss(reduction="none" if self.per else "mean")  # agent.py:317
        self.sync_target(force=True)                                   # agent.py:284

    # dqn/agent.py:80-84 (drives the generator)
    def store(self, obses, actions, rews, dones, new_obses):
        return list(self.replay.store(obses, actions, rews, dones, new_obses))

    # dqn/agent.py:71-78
    @staticmethod
    def _tensorize(rows):
        obs = torch.as_tensor(np.asarray([r[0] for r in rows]), dtype=torch.float32)
        act = torch.as_tensor(np.asarray([r[1] for r in rows]), dtype=torch.int64).unsqueeze(-1)
        rew = torch.as_tensor(np.asarray([r[2] for r in rows]), dtype=torch.float32).unsqueeze(-1)
        done = torch.as_tensor(np.asarray([r[3] for r in rows]), dtype=torch.float32).unsqueeze(-1)
        nxt = torch.as_tensor(np.asarray([r[4] for r in rows]), dtype=torch.float32)
        return obs, act, rew, done, nxt

    def learn(self, u=None, indices=None, trace=None):
        """One replay-minibatch update.  ``u`` (PER) / ``indices`` (uniform) inject the
        sampling randomness; ``trace`` (dict) receives intermediates for parity tests."""
        if self.per:
            w, nodes, rows = self.replay.sample(self.step * self.n_env, u=u)   # agent.py:247
            w_t = torch.as_tensor(np.asarray(w), dtype=torch.float32).unsqueeze(-1)
        else:
            rows = self.replay.sample(indices=indices)                         # agent.py:168/206
            w, nodes, w_t = None, None, None
        obs, act, rew, done, nxt = self._tensorize(rows)

        with torch.no_grad():
            q_next_tgt = self.target(nxt)
            if self.double:                                                    # agent.py:209-214
                q_next_on = self.online(nxt)
                a_star = q_next_on.argmax(dim=1, keepdim=True)
                q_sel = torch.gather(q_next_tgt, 1, a_star)
            else:                                                              # agent.py:171-173
                q_next_on, a_star = None, None
                q_sel = q_next_tgt.max(dim=1, keepdim=True)[0]
            y = rew + (1 - done) * self.gamma * q_sel                          # agent.py:175/216/258

        q = self.online(obs)
        q_sa = torch.gather(q, 1, act)

        abs_td = None
        if self.per:                                                           # agent.py:263-267
            with torch.no_grad():
                abs_td = torch.abs(y - q_sa).detach().cpu().numpy()
                self.replay.write_back(nodes, abs_td)
            loss = torch.mean(w_t * self.huber(q_sa, y))
        else:
            loss = self.huber(q_sa, y)                                         # agent.py:181/222

        self.opt.zero_grad()
        loss.backward()
        if trace is not None:
            trace.update(
                nodes=None if nodes is None else np.asarray(nodes, dtype=np.int64),
                is_w=None if w is None else np.asarray(w, dtype=np.float64),
                obs=obs.numpy().copy(), act=act.numpy().copy(), rew=rew.numpy().copy(),
                done=done.numpy().copy(), nxt=nxt.numpy().copy(),
                q_next_tgt=q_next_tgt.numpy().copy(),
                q_next_on=None if q_next_on is None else q_next_on.numpy().copy(),
                a_star=None if a_star is None else a_star.numpy().copy(),
                y=y.numpy().copy(), q=q.detach().numpy().copy(), q_sa=q_sa.detach().numpy().copy(),
                abs_td=abs_td, loss=float(loss.detach()),
                grads={k: p.grad.detach().numpy().copy() for k, p in self.online.named_parameters()},
            )
        self.opt.step()
        return float(loss.detach())

    # dqn/agent.py:101-110
    def sync_target(self, force=False):
        if (not self.soft and self.step % (self.target_freq // self.n_env) == 0) or force:
            self.target.load_state_dict(self.online.state_dict())
        elif self.soft:
            k = self.tau * self.n_env
            for pt, po in zip(self.target.parameters(), self.online.parameters()):
                pt.data.copy_(k * po.data + (1.0 - k) * pt.data)

    def greedy_actions(self, obses):
        return self.online.greedy(obses)


# --------------------------------------------------------------------------------------
# Pure-numpy restatement of one PER step -- used to check that the
# torch-autograd oracle above and the hand-derived formulas the CUDA kernels implement agree.
# --------------------------------------------------------------------------------------
def numpy_forward(params: dict, x: np.ndarray, dueling: bool):
    """Forward of the macro MLP in float32 numpy; returns (h1, h2, head_out, q)."""
    f = np.float32
    h1 = np.maximum(x @ params["net.0.weight"].T + params["net.0.bias"], f(0))
    h2 = np.maximum(h1 @ params["net.2.weight"].T + params["net.2.bias"], f(0))
    if dueling:
        val = h2 @ params["fc_val.weight"].T + params["fc_val.bias"]
        adv = h2 @ params["fc_adv.weight"].T + params["fc_adv.bias"]
        q = val + (adv - adv.mean(axis=1, keepdims=True, dtype=f))
        return h1, h2, (val, adv), q.astype(f)
    q = h2 @ params["fc_out.weight"].T + params["fc_out.bias"]
    return h1, h2, None, q.astype(f)


def numpy_td_and_grads(online: dict, target: dict, obs, act, rew, done, nxt, is_w, gamma,
                       dueling=True, double=True):
    """Appendix A steps 5-9: TD target, Huber loss and analytic gradients (float32)."""
    f = np.float32
    B = obs.shape[0]
    _, _, _, q_next_tgt = numpy_forward(target, nxt, dueling)
    if double:
        _, _, _, q_next_on = numpy_forward(online, nxt, dueling)
        a_star = q_next_on.argmax(axis=1)
        q_sel = q_next_tgt[np.arange(B), a_star]
    else:
        q_sel = q_next_tgt.max(axis=1)
    y = rew.reshape(-1) + ((f(1) - done.reshape(-1)) * f(gamma)) * q_sel
    h1, h2, _, q = numpy_forward(online, obs, dueling)
    a = act.reshape(-1)
    q_sa = q[np.arange(B), a]
    delta = q_sa - y
    w = np.ones(B, f) if is_w is None else is_w.reshape(-1).astype(f)
    ell = np.where(np.abs(delta) < 1, f(0.5) * delta * delta, np.abs(delta) - f(0.5)).astype(f)
    loss = f((w * ell).sum(dtype=f) / f(B))
    g = (w / f(B)) * np.clip(delta, -1, 1)
    A = q.shape[1]
    dq = np.zeros((B, A), f)
    dq[np.arange(B), a] = g
    grads = {}
    if dueling:
        dval = dq.sum(axis=1, keepdims=True)
        dadv = dq - dq.mean(axis=1, keepdims=True)
        grads["fc_val.weight"] = dval.T @ h2
        grads["fc_val.bias"] = dval.sum(axis=0)
        grads["fc_adv.weight"] = dadv.T @ h2
        grads["fc_adv.bias"] = dadv.sum(axis=0)
        dh2 = dval @ online["fc_val.weight"] + dadv @ online["fc_adv.weight"]
    else:
        grads["fc_out.weight"] = dq.T @ h2
        grads["fc_out.bias"] = dq.sum(axis=0)
        dh2 = dq @ online["fc_out.weight"]
    dz2 = dh2 * (h2 > 0)
    grads["net.2.weight"] = dz2.T @ h1
    grads["net.2.bias"] = dz2.sum(axis=0)
    dz1 = (dz2 @ online["net.2.weight"]) * (h1 > 0)
    grads["net.0.weight"] = dz1.T @ obs
    grads["net.0.bias"] = dz1.sum(axis=0)
    return dict(y=y.astype(f), q=q, q_sa=q_sa, abs_td=np.abs(y - q_sa).astype(f), loss=loss,
                grads={k: v.astype(f) for k, v in grads.items()})


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (exact product, one rounding)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def numpy_adam(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor step with the rounding sequence torch 2.11's CPU kernels
    execute (Appendix A step 11; found by bit-matching, see test_numpy_adam_bit_matches_torch):
    m <- fma(1-b1, g-m, m) [lerp_]; v <- fma((1-b2) g, g, b2 v) [mul_ + addcmul_]; python-float64
    bias corrections; denom = sqrt(v)/sqrt(bc2) + eps; p <- p + (-(lr/bc1) m) / denom [addcdiv_]."""
    f = np.float32
    m = _fma32(f(1 - b1), (g - m).astype(f), m)
    v = _fma32((f(1 - b2) * g).astype(f), g, (v * f(b2)).astype(f))
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    step_size = lr / bc1
    denom = ((np.sqrt(v).astype(f) / f(bc2 ** 0.5)).astype(f) + f(eps)).astype(f)
    p = (p + ((f(-step_size) * m).astype(f) / denom).astype(f)).astype(f)
    return p, m, v


def synthetic_transitions(n: int, obs_dim: int, seed: int = 20251018, n_actions: int = 8):
    """U[0,1) float32 states chained s'_t = s_{t+1}, last feature from the
    8-level action grid, actions U{0..A-1}, rewards clip(N(0.3,1.5^2),-24,3), done every 90th."""
    rng = np.random.default_rng(seed)
    s = rng.random((n + 1, obs_dim), dtype=np.float32)
    s[:, -1] = (rng.integers(1, n_actions + 1, size=n + 1) / n_actions).astype(np.float32)
    a = rng.integers(0, n_actions, size=n).astype(np.int64)
    r = np.clip(rng.normal(0.3, 1.5, size=n), -24.0, 3.0).astype(np.float32)
    d = np.zeros(n, dtype=np.float32)
    d[89::90] = 1.0
    return s[:-1].copy(), a, r, d, s[1:].copy()
