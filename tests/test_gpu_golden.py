"""CUDA engine (through the C ABI) replaying the golden trajectories recorded from the unmodified
reference: every output, the observation digest and the board-state digest of every step."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import golden_util as GU

pytestmark = pytest.mark.gpu


def _state_dict(eng, rec, base_none):
    st = eng.decode_state(rec)
    h = st["header"]
    return dict(cost_def=float(h["cost_def"]), cost_atk=float(h["cost_atk"]),
                base_LP=None if base_none else int(h["base_LP"]), steps=int(h["steps"]),
                attacker_cd=int(h["attacker_cd"]), defender_cd=int(h["defender_cd"]), map6=st["map6"],
                towers=[(int(r["loc"]), int(r["type_lv"]) & 3, int(r["type_lv"]) >> 2, float(r["cd"])) for r in st["towers"]],
                enemies=[(int(r["loc"]), int(r["type_lv"]) & 3, float(r["LP"]), 0.0, float(r["margin"]), 0,
                          int(r["slowdown"])) for r in st["enemies"]])


def _replayable(path):
    return True


@pytest.mark.parametrize("path", [p for p in GU.trajectories() if _replayable(p)],
                         ids=lambda p: os.path.basename(p)[5:-4])
def test_cuda_replays_reference_trajectory(path):
    import torch
    from gym_td_b200 import engine as E
    from tests import parity_util as PU
    traj = GU.Traj(path)
    z, L, kind = traj.z, traj.L, traj.kind
    cfg = PU.make_config(**traj.config_overrides())
    base_none = "base_LP" in traj.meta["overrides"] and traj.meta["overrides"]["base_LP"] is None
    eng = E.Engine(kind, L, 1, cfg=cfg)
    eng.upload_maps([E.map_from_planes(L, int(z["num_roads"]), [int(x) for x in z["start"]], int(z["end"]),
                                       z["road"], z["dist"], z["dir"])])
    dev = torch.device("cuda", 0)
    use_np = not traj.meta["random_agent"]
    np_rs = None
    if kind != "2p" and not use_np:
        eng.seed_opponent(z["py_state"].reshape(1, 625))
        eng.set_difficulty(traj.meta["difficulty"])
    if use_np:
        np_rs = np.random.RandomState()
        np_rs.set_state(("MT19937", z["np_state"][:624], int(z["np_state"][624]), 0, 0.0))
    obs = torch.empty((1, 45, L, L), dtype=torch.float32, device=dev)
    eng.reset(obs=obs)
    torch.cuda.synchronize()
    assert GU.digest64(obs[0].cpu().numpy().tobytes()) == z["obs_digest"][0]
    out = dict(reward=torch.zeros(1, dtype=torch.float64, device=dev), done=torch.zeros(1, dtype=torch.uint8, device=dev),
               win=torch.zeros(1, dtype=torch.int8, device=dev), allow=torch.zeros(1, dtype=torch.uint8, device=dev),
               fail_def=torch.zeros(1, dtype=torch.int32, device=dev), fail_atk=torch.zeros((1, 4), dtype=torch.int32, device=dev),
               real_atk=torch.zeros((1, 3, 8), dtype=torch.int64, device=dev))
    real_def = torch.zeros((1, 6, L, L) if traj.multi else (1,), dtype=torch.int64, device=dev)
    atk_cd = def_cd = 0
    diff = traj.meta["difficulty"]
    for t in range(1, traj.T + 1):
        d = a = opp = cluster = None
        if kind != "atk":
            d = torch.from_numpy(np.ascontiguousarray(traj.def_action(t)).reshape(real_def.shape)).to(dev)
        if kind != "def":
            a = torch.from_numpy(traj.atk_action(t).reshape(1, 3, 8)).to(dev)
        if use_np and kind == "def":     # host-resolved scripted attacker on the env's np_random stream
            atk_cd = max(atk_cd - 1, 0)
            byte, word = 0xFF, 0xFFFFFFFF
            if atk_cd == 0:
                if diff == 1:                                           # TDGymBasic.py:102-103
                    tt = int(np_rs.randint(0, 4))
                    rd = int(np_rs.randint(int(z["num_roads"])))
                    byte = tt | (rd << 4)
                else:                                                   # TDGymBasic.py:87-89
                    cl = np_rs.randint(0, 4, [8], dtype=np.int64)
                    rd = int(np_rs.randint(int(z["num_roads"])))
                    word = sum(int(c) << (2 * k) for k, c in enumerate(cl)) | (rd << 16)
                atk_cd = cfg.attacker_action_interval
            if diff == 1:
                opp = torch.tensor([byte], dtype=torch.uint8, device=dev)
            else:
                cluster = torch.tensor([word], dtype=torch.int64, device=dev).to(torch.uint32)
        if use_np and kind == "atk":     # random_tower_lv0 on np_random (TDGymBasic.py:112,118-121)
            assert diff == 0
            build = -1
            if max(def_cd - 1, 0) == 0:
                r, c = np_rs.randint(0, L, [2, ])
                tt = int(np_rs.randint(0, 4))
                build = tt * L * L + int(r) * L + int(c)
            d = torch.tensor([build], dtype=torch.int64, device=dev)
        io = E.Engine.make_io(def_action=d, atk_action=a, opponent=opp, opponent_cluster=cluster,
                              multi_action=traj.multi, obs=obs,
                              reward=out["reward"], done=out["done"], win=out["win"], allow_next=out["allow"],
                              real_def=real_def, real_atk=out["real_atk"], fail_def=out["fail_def"],
                              fail_atk=out["fail_atk"])
        eng.step(io, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        GU.check_outputs(traj, t, float(out["reward"][0]), int(out["done"][0]), int(out["win"][0]),
                         int(out["allow"][0]), None if traj.multi else int(real_def[0]), int(out["fail_def"][0]),
                         out["real_atk"][0].cpu().numpy(), out["fail_atk"][0].tolist(),
                         real_def[0].cpu().numpy() if traj.multi else None)
        assert GU.digest64(obs[0].cpu().numpy().tobytes()) == z["obs_digest"][t], "%s obs step %d" % (traj.name, t)
        sd = _state_dict(eng, eng.get_state_raw(0, 1)[0], base_none)
        def_cd = sd["defender_cd"]
        assert GU.digest64(GU.state_bytes(sd)) == z["state_digest"][t - 1], "%s state step %d" % (traj.name, t)
    eng.close()
