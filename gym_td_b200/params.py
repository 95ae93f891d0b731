"""Game configuration: same names and behaviour as the reference's gym_TD/envs/TDParam.py.

`config` is a mutable object changed through `paramConfig(**kwargs)` (a bare setattr per key,
TDParam.py:98-100); `hyper_parameters` refuses attribute assignment (TDParam.py:112-113).
The engine snapshots `config` into a `td_config` struct when an env is created or reset.
"""


class Config(object):
    # values: gym_TD/envs/TDParam.py:2-94
    def __init__(self):
        self.max_enemy_lv = 1
        self.max_tower_lv = 1
        self.enemy_types = 4
        self.tower_types = 4
        self.enemy_LP = [[820, 1700], [2050, 3000], [6000, 8000], [8000, 12000]]
        self.enemy_speed = [[.25, .25], [.13, .13], [.1, .1], [.1, .1]]
        self.enemy_defense = [[0, 0], [200, 250], [600, 800], [80, 100]]
        self.enemy_cost = [[8, 8], [15, 15], [40, 40], [30, 30]]
        self.tower_attack = [[454, 540], [651, 771], [566, 691], [358, 424]]
        self.tower_range = [[3, 3], [2, 2], [4, 4], [3, 3]]
        self.tower_splash_range = [[0, 0], [0, 0], [1, 1], [0, 0]]
        self.tower_cost = [[10, 10], [17, 17], [23, 23], [12, 12]]
        self.tower_attack_interval = [[2, 2], [4, 4], [7, 7], [4.75, 4.75]]
        self.tower_destruct_return = .5
        self.frozen_time = 2
        self.frozen_ratio = .2
        self.attacker_init_cost = 0
        self.defender_init_cost = 10
        self.base_LP = 5
        self.max_cost = 100
        self.reward_kill = 0.1
        self.penalty_leak = 10.
        self.reward_time = 0.001
        self.attacker_cost_init_rate = .5
        self.attacker_cost_final_rate = 1
        self.defender_cost_rate = .2
        self.tower_distance = 2
        self.enemy_upgrade_at = 0.75
        self.attacker_action_interval = 1
        self.defender_action_interval = 1


config = Config()


def paramConfig(**kwargs):
    for key, val in kwargs.items():
        setattr(config, key, val)


def getConfig():
    return config.__dict__


class HyperParameters(object):
    # gym_TD/envs/TDParam.py:105-113
    def __init__(self):
        d = super(HyperParameters, self).__setattr__
        d('max_episode_steps', 1200)
        d('video_frames_per_second', 50)
        d('allow_multiple_actions', False)
        d('max_cluster_length', 8)
        d('max_num_of_roads', 3)

    def __setattr__(self, name, value):
        raise RuntimeError('You are not supposed to modify hyper parameters during runtime.')


hyper_parameters = HyperParameters()


def getHyperParameters():
    return hyper_parameters.__dict__.copy()


def n_channels():
    """TDBoard.n_channels() (TDBoard.py:146-154)."""
    return 15 + 2 * config.tower_types + config.max_tower_lv + 1 + 5 * config.enemy_types
