"""Defaults of the reference's configuration for this path, restated as data.

``HYPER_PARAMS``: env/dqn_config.py:26-56.  ``network_config``: the macro-state MLP of
env/custom_env/macro with lane/dqn_config.py:58-104 (D -> 256 -> ReLU -> 128 -> ReLU, fc_out_dim 128,
Adam, SmoothL1).  The reference files themselves import the SUMO bindings and cannot be imported
where SUMO is absent, hence the restatement.
"""
import torch.nn as nn
import torch.optim as optim

HYPER_PARAMS = {
    "gpu": "0", "n_env": 1, "lr": 1e-4, "gamma": 0.99, "eps_start": 1.0, "eps_min": 0.01, "eps_dec": 2e6,
    "eps_dec_exp": True, "bs": 32, "min_mem": 100000, "max_mem": 1000000, "target_update_freq": 30000,
    "target_soft_update": True, "target_soft_update_tau": 1e-3, "save_freq": 10000, "log_freq": 4500,
    "algo": "DuelingDoubleDQNAgent",
}


class ObsSpace:
    """Stand-in for the gym Box the reference passes as ``input_dim`` (only ``.shape`` is read)."""

    def __init__(self, dim):
        self.shape = (int(dim),)


def network_config(input_dim_space):
    d = input_dim_space.shape[0]
    net = nn.Sequential(nn.Linear(d, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU())
    return net, 128, optim.Adam, nn.SmoothL1Loss


def network_config_elu(input_dim_space):
    """The same body with the activation of the repo-HEAD config (``ACTIVATION = nn.ELU()``, env/dqn_config.py:175)."""
    d = input_dim_space.shape[0]
    net = nn.Sequential(nn.Linear(d, 256), nn.ELU(), nn.Linear(256, 128), nn.ELU())
    return net, 128, optim.Adam, nn.SmoothL1Loss


class TwoStreamBody(nn.Module):
    """The repo-HEAD body (env/dqn_config.py:66-193): state = [macro | grid C,H,W]; grid -> (Conv2d 3x3 p1 + act) x n ->
    flatten; cat([flatten, macro]) -> (Linear + act) x m.  Attribute names (``cnn_stream``, ``dense_stream``,
    ``macro_len``, ``micro_shape``) are the ones the reference's checkpoints and ``network.match_hybrid_body`` expect.
    The module only carries parameters and structure: compute runs in librmc_b200."""

    def __init__(self, macro_len=14, micro_shape=(2, 27, 5), cnn=((32, (1, 1)), (64, (2, 1)), (64, (2, 2))), dense=(512, 256),
                 activation=nn.ELU):
        super().__init__()
        act = activation()
        self.macro_len, self.micro_shape = int(macro_len), tuple(int(v) for v in micro_shape)
        layers, ch, h, w = [], self.micro_shape[0], self.micro_shape[1], self.micro_shape[2]
        for out_ch, (sh, sw) in cnn:
            layers += [nn.Conv2d(ch, out_ch, kernel_size=(3, 3), stride=(sh, sw), padding=(1, 1)), act]
            ch, h, w = out_ch, (h + 2 - 3) // sh + 1, (w + 2 - 3) // sw + 1
        self.cnn_stream = nn.Sequential(*layers)
        layers, width = [], ch * h * w + self.macro_len
        for out_f in dense:
            layers += [nn.Linear(width, out_f), act]
            width = out_f
        self.dense_stream = nn.Sequential(*layers)
        self.fc_out_dim = width


def network_config_hybrid(input_dim_space):
    """``network_config`` of env/dqn_config.py:148-193 (macro 14 + grid 2x27x5, CNN 32/64/64, dense 512/256, ELU)."""
    net = TwoStreamBody()
    return net, net.fc_out_dim, optim.Adam, nn.SmoothL1Loss


HYBRID_OBS_DIM = 14 + 2 * 27 * 5


def make_agent(algo, obs_dim, batch_size, buffer_size, *, save_dir, log_dir, n_actions=8, gpu="0", activation="relu",
               **overrides):
    """Construct an agent the way train.py:24-48 does, with the reference defaults."""
    from . import agent as Agents
    hp = dict(HYPER_PARAMS)
    hp.update(overrides)
    cls = getattr(Agents, algo)
    return cls(n_env=hp["n_env"], lr=hp["lr"], gamma=hp["gamma"], epsilon_start=hp["eps_start"],
               epsilon_min=hp["eps_min"], epsilon_decay=hp["eps_dec"], epsilon_exp_decay=hp["eps_dec_exp"],
               nn_conf_func=(network_config_hybrid if activation == "hybrid" else network_config_elu if activation == "elu" else network_config),
               input_dim=ObsSpace(obs_dim),
               output_dim=n_actions,
               batch_size=batch_size, min_buffer_size=min(hp["min_mem"], buffer_size), buffer_size=buffer_size,
               update_target_frequency=hp["target_update_freq"], target_soft_update=hp["target_soft_update"],
               target_soft_update_tau=hp["target_soft_update_tau"], save_frequency=hp["save_freq"],
               log_frequency=hp["log_freq"], save_dir=save_dir, log_dir=log_dir, load=False, algo=algo, gpu=gpu)
