"""Bring-up diagnostics run on the GPU box: each stage dumps raw kernel outputs next to oracle values into
gpurun_out/diag_<stage>.npz so that a failure can be analysed offline (no GPU in the build container).

    python tools/diag.py umma | fwd | bwd | sim
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
OUT = REPO / "gpurun_out"
OUT.mkdir(exist_ok=True)

import spn4cir_b200 as spn  # noqa: E402
from spn4cir_b200 import _lib, synth  # noqa: E402


def umma():
    res = {}
    for mode in (0, 1, 2, 3):
        for (N, K) in ((128, 64), (128, 128), (64, 128), (128, 256)):
            if (mode & 2) and K != 128:
                continue
            g = torch.Generator().manual_seed(100 * mode + N + K)
            a = torch.randn((128, K), generator=g).to(torch.bfloat16)
            b = (torch.randn((K, N), generator=g) if mode & 2 else torch.randn((N, K), generator=g)).to(torch.bfloat16)
            ref = a.float() @ (b.float() if mode & 2 else b.float().T)
            # probe: A[m, k] = 1 iff k == m % K ; B unique small integers -> C[m, n] = B(n, m % K)
            ap = torch.zeros((128, K))
            ap[torch.arange(128), torch.arange(128) % K] = 1.0
            if mode & 2:
                bp = (torch.arange(K)[:, None] * 1.0 + torch.arange(N)[None, :] / 128.0)
            else:
                bp = (torch.arange(K)[None, :] * 1.0 + torch.arange(N)[:, None] / 128.0)
            outs = []
            for (aa, bb) in ((a, b), (ap.to(torch.bfloat16), bp.to(torch.bfloat16))):
                ad, bd = aa.cuda().contiguous(), bb.cuda().contiguous()
                c = torch.full((128, N), float("nan"), device="cuda")
                _lib.call("spn_debug_umma_gemm", ad.data_ptr(), bd.data_ptr(), c.data_ptr(), N, K, mode,
                          torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                outs.append(c.cpu().numpy())
            err = np.abs(outs[0] - ref.numpy()).max()
            refp = ap.to(torch.bfloat16).float() @ (bp.to(torch.bfloat16).float() if mode & 2 else bp.to(torch.bfloat16).float().T)
            errp = np.abs(outs[1] - refp.numpy()).max()
            print(f"umma mode={mode} N={N} K={K}: random max|err|={err:.4g} probe max|err|={errp:.4g}", flush=True)
            res[f"m{mode}_N{N}_K{K}_rand"] = outs[0]
            res[f"m{mode}_N{N}_K{K}_probe"] = outs[1]
    np.savez_compressed(OUT / "diag_umma.npz", **res)


def fwd():
    from oracle import bank_ce as obank
    res = {}
    for (B, N, D, T, seed) in ((16, 333, 128, 1, 1), (64, 3000, 640, 1, 5), (6, 50, 256, 32, 11)):
        q, bank, labels = synth.make_bank_case(B, N, D, T, "S", seed)
        tau = 0.02 if T == 1 else 0.05
        db = spn.DeviceBank.from_tensor(bank, "cuda")
        st = spn.bank_ce_stats(q.cuda(), db, labels.cuda(), tau, want_ssum=True)
        torch.cuda.synchronize()
        qb, bb = q.to(torch.bfloat16).float(), bank.to(torch.bfloat16).float()
        if T == 1:
            m, l, s_t, ss = obank.shard_stats(qb, bb, labels, tau)
        else:
            s = torch.einsum("bd,ntd->bnt", qb.double(), bb.double()).amax(-1) / tau
            m = s.max(1).values
            l = torch.exp(s - m[:, None]).sum(1)
            s_t = s[torch.arange(B), labels]
            ss = s.sum(1)
        got = st.cpu().double()
        lse_got, lse_ref = got[0] + torch.log(got[1]), m + torch.log(l)
        print(f"fwd B={B} N={N} D={D} T={T}: max|lse err|={(lse_got - lse_ref).abs().max():.4g} "
              f"max|m err|={(got[0] - m).abs().max():.4g} max|st err|={(got[2] - s_t).abs().max():.4g} "
              f"max|ssum rel|={((got[3] - ss).abs() / ss.abs().clamp_min(1)).max():.4g}", flush=True)
        res[f"B{B}_N{N}_D{D}_T{T}_got"] = got.numpy()
        res[f"B{B}_N{N}_D{D}_T{T}_ref"] = torch.stack([m, l, s_t, ss]).numpy()
    np.savez_compressed(OUT / "diag_fwd.npz", **res)


def bwd():
    from oracle import bank_ce as obank
    res = {}
    for (B, N, D, T, seed) in ((16, 333, 128, 1, 1), (64, 3000, 640, 1, 5), (64, 2500, 768, 1, 6), (6, 50, 256, 32, 11)):
        q, bank, labels = synth.make_bank_case(B, N, D, T, "S", seed)
        tau = 0.02 if T == 1 else 0.05
        qd = q.cuda().requires_grad_(True)
        loss = spn.bank_infonce(qd, spn.DeviceBank.from_tensor(bank, "cuda"), labels.cuda(), tau)
        loss.backward()
        torch.cuda.synchronize()
        qb, bb = q.to(torch.bfloat16).float(), bank.to(torch.bfloat16).float()
        if T == 1:
            o_loss, o_dq, _ = obank.infonce_loss_and_grads(qb, bb, labels, tau, dtype=torch.float64)
        else:
            o_loss, o_dq, _ = obank.blip2_loss_and_grads(qb, bb, labels, tau, dtype=torch.float64)
        dq = qd.grad.cpu().double()
        rel = float((dq - o_dq).norm() / o_dq.norm())
        print(f"bwd B={B} N={N} D={D} T={T}: loss={loss.item():.6f} ref={o_loss.item():.6f} dq rel={rel:.4g} "
              f"|dq|={dq.norm():.4g} |ref|={o_dq.norm():.4g}", flush=True)
        res[f"B{B}_N{N}_D{D}_T{T}_dq"] = dq.numpy()
        res[f"B{B}_N{N}_D{D}_T{T}_ref"] = o_dq.numpy()
    np.savez_compressed(OUT / "diag_bwd.npz", **res)


def sim():
    from oracle import retrieval as oretr
    q, gal, _, ref, tgt, grp = synth.make_retrieval_case(70, 1000, 128, 1, 10, noise=6.0)
    s = q @ gal.T
    v, i = spn.sim_topk(q.cuda(), gal.cuda(), 50, exclude=ref)
    r, g = spn.sim_rank(q.cuda(), gal.cuda(), tgt, exclude=ref, group=grp)
    torch.cuda.synchronize()
    wv, wi = oretr.topk(s, 50, exclude=ref.tolist())
    wr, wg = oretr.rank_of_target(s, tgt.tolist(), ref.tolist(), grp.tolist())
    print(f"sim: topk idx mismatch={(i.cpu().numpy() != wi).mean():.4g} val err={np.abs(v.cpu().numpy() - wv).max():.4g} "
          f"rank mismatch={(r.cpu().numpy() != wr).mean():.4g} grank mismatch={(g.cpu().numpy() != wg).mean():.4g}", flush=True)
    np.savez_compressed(OUT / "diag_sim.npz", v=v.cpu().numpy(), i=i.cpu().numpy(), wv=wv, wi=wi, r=r.cpu().numpy(), wr=wr)


if __name__ == "__main__":
    {"umma": umma, "fwd": fwd, "bwd": bwd, "sim": sim}[sys.argv[1]]()
