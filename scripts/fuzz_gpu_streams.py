"""GPU fuzz of concurrent use: the loss launched on TWO non-default streams back to back without synchronisation (different random
shapes / modes / inputs, also DS / DC with their side-stream instance-mask preparation), repeated; every result must equal the
same call made alone on the default stream, bit for bit.  python scripts/fuzz_gpu_streams.py [N] [seed0]"""
import os
import random
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from mdn_sfm_b200 import synthetic
from mdn_sfm_b200.loss_functions import Loss

DEV = "cuda"


def run(loss, batch, scales):
    inputs, flows, mobiles, cams, inst = batch
    f = {k: v.clone().requires_grad_(True) for k, v in flows.items()}
    m = {k: v.clone().requires_grad_(True) for k, v in mobiles.items()}
    c = {k: v.clone().requires_grad_(True) for k, v in cams.items()}
    _, losses = loss(inputs, [-1, 1], f, m, inst, list(scales), c)
    losses["loss"].backward()
    return [losses["loss"].detach()] + [t.grad for d in (f, m, c) for t in d.values()]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(n):
        jobs = []
        for _ in range(2):
            B, H, W = rng.randint(1, 3), 8 * rng.randint(2, 14), 8 * rng.randint(4, 34)
            scales = (0, 1, 2, 3)[:rng.randint(1, 3)]
            mode = rng.choice(["SN", "T", "TG", "DS", "DC"])
            opt = synthetic.default_opt(B, H, W, scales=list(scales), threshold=0.8625 if mode == "TG" else 9.22)
            batch = synthetic.make_batch(B, H, W, scales=scales, seed=rng.randint(0, 10 ** 6), flow_std=0.05, with_instances=mode in ("DS", "DC"), device=DEV)
            if batch[4] is not None:
                batch = batch[:4] + ([{"instances": d["instances"].to(DEV)} for d in batch[4]],)
            jobs.append(dict(B=B, H=H, W=W, scales=scales, mode=mode, batch=batch, loss=Loss(opt, no_ssim=False, mode=mode, photometric=True)))
        case = [{k: j[k] for k in ("B", "H", "W", "scales", "mode")} for j in jobs]
        try:
            alone = [run(j["loss"], j["batch"], j["scales"]) for j in jobs]
            torch.cuda.synchronize()
            for rep in range(3):
                outs = [None, None]
                for k in (0, 1):
                    streams[k].wait_stream(torch.cuda.current_stream())
                for k in (0, 1):
                    with torch.cuda.stream(streams[k]):
                        outs[k] = run(jobs[k]["loss"], jobs[k]["batch"], jobs[k]["scales"])
                torch.cuda.synchronize()
                for k in (0, 1):
                    for a, b in zip(alone[k], outs[k]):
                        assert torch.equal(a, b), ("stream", k, "rep", rep)
        except AssertionError as e:
            fails += 1
            print("FAIL", it, case, "->", str(e)[:300], flush=True)
        except Exception as e:
            fails += 1
            print("ERROR", it, case, "->", type(e).__name__, str(e)[:300], flush=True)
            traceback.print_exc(limit=5)
    print("fuzz (streams): %d pairs of concurrent losses x 3 repetitions, %d failures" % (n, fails))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
