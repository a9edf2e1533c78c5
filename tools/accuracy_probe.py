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
    net.target_changed()                     # a direct write to the target parameters: re-derive their operand planes
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
        if os.environ.get("PROBE_DIAG"):
            tr = net.online_s.trunk
            g3 = store.export_named(store.grad)
            kname, bname = agent.net_def.trunk.names[2]
            dz3 = tr.dz_planes[2].to_dense().double()                     # [49 * B, 64], rows q * B + b
            act2 = tr.act_planes[1].to_dense().double()                   # [81 * B, 64]
            want_b = dz3.sum(0).cpu().numpy()
            print("   [diag] conv3 bias grad vs column sum of the device's dz3: max abs %.3e (scale %.3e); vs fp64 oracle %.3e"
                  % (np.abs(g3[bname] - want_b).max(), np.abs(want_b).max(),
                     np.abs(g3[bname] - r64["grads"][bname].numpy()).max()))
            a2 = act2.reshape(9, 9, B, 64)
            d3 = dz3.reshape(7, 7, B, 64)
            want_w = torch.zeros(3, 3, 64, 64, dtype=torch.float64, device="cuda")
            for ky in range(3):
                for kx in range(3):
                    xs = a2[ky:ky + 7, kx:kx + 7].reshape(-1, 64)
                    want_w[ky, kx] = xs.t() @ d3.reshape(-1, 64)
            want_w = want_w.cpu().numpy()
            print("   [diag] conv3 kernel grad vs act2^T dz3 from the device's planes: max abs %.3e (scale %.3e); "
                  "that recomputation vs fp64 oracle %.3e"
                  % (np.abs(g3[kname] - want_w).max(), np.abs(want_w).max(),
                     np.abs(want_w - r64["grads"][kname].numpy()).max()))
            dzf = tr.dz_planes[3].to_dense().double()
            print("   [diag] fc1 dz planes vs fp32 dz: max abs %.3e" % (dzf - tr.dzs[3].double()).abs().max().item())
        if os.environ.get("PROBE_DIAG2"):
            import torch.nn.functional as F
            tr = net.online_s.trunk
            P = [torch.from_numpy(v).double().requires_grad_(True) for v in online.values()]
            x = torch.from_numpy(ob["states"]).double() / 255.0
            hcur = x.permute(0, 3, 1, 2)
            zs, k = [], 0
            for stride in (4, 2, 1):
                z = F.conv2d(hcur, P[k].permute(3, 2, 0, 1), P[k + 1], stride=stride)
                z.retain_grad()
                zs.append(z)
                hcur = F.relu(z)
                k += 2
            flat = hcur.permute(0, 2, 3, 1).reshape(B, -1)
            z4 = flat @ P[k] + P[k + 1]
            z4.retain_grad()
            zs.append(z4)
            qq = F.relu(z4) @ P[k + 2] + P[k + 3]
            loss64 = on.q_head_loss(qq, torch.from_numpy(r64["targets"]).double(),
                                    torch.from_numpy(ob["weights"].astype(np.float64)), True)
            loss64.backward()
            for i, z in enumerate(zs[:3]):
                npix = z.shape[2] * z.shape[3]
                act_ref = F.relu(z).detach().permute(2, 3, 0, 1).reshape(npix * B, -1)      # rows pixel * B + b
                dz_ref = z.grad.permute(2, 3, 0, 1).reshape(npix * B, -1)
                act_dev = tr.act_planes[i].to_dense().double().cpu()
                dz_dev = tr.dz_planes[i].to_dense().double().cpu()
                print("   [diag2] conv%d: act max abs err %.2e (scale %.2e); dz max abs err %.2e (scale %.2e); "
                      "mask mismatches %d" % (i + 1, (act_dev - act_ref).abs().max(), act_ref.abs().max(),
                                               (dz_dev - dz_ref).abs().max(), dz_ref.abs().max(),
                                               int(((act_dev > 0) != (act_ref > 0)).sum())))
            dzf_dev = tr.dz_planes[3].to_dense().double().cpu()
            print("   [diag2] fc1: dz max abs err %.2e (scale %.2e)" % ((dzf_dev - zs[3].grad).abs().max(),
                                                                       zs[3].grad.abs().max()))
        got = store.export_named(store.grad)
        for name in r64["grads"]:
            w = r64["grads"][name].numpy()
            s = np.abs(w).max() + 1e-30
            print("   %-45s ours %.2e  fp32 oracle %.2e   (max abs err / max |grad|)"
                  % (name.split("network_0/")[-1], np.abs(got[name] - w).max() / s,
                     np.abs(r32["grads"][name].numpy() - w).max() / s))

if __name__ == "__main__":
    main()
