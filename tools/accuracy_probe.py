"""Relative errors of one B = 512 DQN + PER learn step against the fp64 evaluation of the same graph (and the fp32
oracle's own distance to it), for a given accumulation cap:  CB200_TILED_MAX_CHUNKS=10 python tools/accuracy_probe.py"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import nets as on            # noqa: E402
import test_learn_gpu as T               # noqa: E402


def main():
    B, A = 512, 6
    agent = T._make_agent((84, 84, 4), A, B, False, False, True, None, True)
    rng = np.random.RandomState(3)
    n = 1024
    agent.memory.store_columns({"state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
                                "next_state:observation": rng.randint(0, 256, (n, 84, 84, 4)).astype(np.uint8),
                                "action": rng.randint(0, A, n).astype(np.int64),
                                "reward": rng.randint(-1, 2, n).astype(np.float64),
                                "game_over": (rng.rand(n) < 0.1).astype(np.uint8)})
    agent.memory.update_priorities(np.arange(n), np.abs(rng.randn(n)))
    store, net = agent.net_def.store, agent.networks["main"]
    if "PROBE_LR" in os.environ:
        net.params.learning_rate = float(os.environ["PROBE_LR"])
    net.theta_target.copy_(store.theta * 0.9 + 0.01)
    for step in range(int(os.environ.get("PROBE_STEPS", "2"))):
        online, target = store.export_named(), store.export_named(net.theta_target)
        random.seed(10 + step)
        np.random.seed(10 + step)
        batch = agent.sample_batch()
        loss, _, gnorm = agent.learn_from_batch(batch)
        torch.cuda.synchronize()
        for k in ("state:observation", "next_state:observation"):
            batch.column(k)
        cols = {k: v.cpu().numpy() for k, v in batch.columns.items()}
        ob = dict(states=cols["state:observation"], next_states=cols["next_state:observation"], actions=cols["action"],
                  rewards=cols["reward"], game_overs=cols["game_over"].astype(bool), weights=cols["weight32"])
        res = {}
        for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            o = on.QNetOracle((84, 84, 4), A, False, dt)
            opt = on.AdamTF([torch.from_numpy(v).to(dt) for v in online.values()], float(net.params.learning_rate), 0.9, 0.99, 1e-4, dtype=dt)
            res[name] = on.dqn_learn_step(o, o.cast(online), o.cast(target), opt, ob, 0.99, True, False, None)
        r64, r32 = res["f64"], res["f32"]
        print("step %d cap %s: loss rel err ours %.2e (fp32 oracle %.2e); grad_norm rel err ours %.2e (fp32 oracle %.2e); "
              "q_online max err %.2e" % (step, os.environ.get("CB200_TILED_MAX_CHUNKS", "20"),
                                         abs(loss - r64["loss"]) / r64["loss"],
                                         abs(r32["loss"] - r64["loss"]) / r64["loss"],
                                         abs(gnorm - r64["grad_norm"]) / r64["grad_norm"],
                                         abs(r32["grad_norm"] - r64["grad_norm"]) / r64["grad_norm"],
                                         np.abs(net.online_s.q.cpu().numpy() - r64["q_online"]).max()))
        if os.environ.get("PROBE_DIAG"):
            tr = net.online_s.trunk
            fc1, conv3 = tr.layers[3], tr.layers[2]
            W = store.view(store.theta, agent.net_def.trunk.names[3][0]).double()          # [3136, 512] AFTER Adam!
            W_before = torch.from_numpy(online[agent.net_def.trunk.names[3][0]]).cuda().double()
            npix, Ca, N = 49, 64, 512
            for tag, Wm in (("theta before this step's Adam", W_before), ("theta after Adam", W)):
                want_wT = Wm.reshape(npix, Ca, N).permute(0, 2, 1).reshape(npix * N, Ca)
                got_wT = fc1.wT_planes.to_dense().double()
                print("   [diag] fc1 wT planes vs %s: max abs diff %.3e" % (tag, (got_wT - want_wT).abs().max().item()))
            dzf = tr.dz_planes[3].to_dense().double() if tr.dz_planes[3] is not None else tr.dzs[3].double()
            act3 = tr.act_planes[2].to_dense().double()                                      # [49 * B, 64]
            want = torch.zeros(npix * B, Ca, dtype=torch.float64, device="cuda")
            for q in range(npix):
                want[q * B:(q + 1) * B] = dzf @ W_before[q * Ca:(q + 1) * Ca].t()
            want = want * (act3 > 0)
            got_dz3 = tr.dz_planes[2].to_dense().double()
            print("   [diag] dz3 (fc1.bwd_x output) vs recomputation with pre-Adam weights: max abs %.3e / scale %.3e"
                  % ((got_dz3 - want).abs().max().item(), want.abs().max().item()))
        got = store.export_named(store.grad)
        for name in r64["grads"]:
            w = r64["grads"][name].numpy()
            s = np.abs(w).max() + 1e-30
            print("   %-45s ours %.2e  fp32 oracle %.2e   (max abs err / max |grad|)"
                  % (name.split("network_0/")[-1], np.abs(got[name] - w).max() / s,
                     np.abs(r32["grads"][name].numpy() - w).max() / s))

if __name__ == "__main__":
    main()
