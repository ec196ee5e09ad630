"""Multi-seed PSNR twin (the second half of BASELINE.json's metric): the UNMODIFIED run_nerf.py train() on one synthetic
Blender-format scene (BASELINE config 2 shape: 3 training views, 400 x 400, 64 + 128 samples, N_rand 4096, white background),
reference GPU eager vs this package through the drop-in, several seeds per arm so that the seed-to-seed spread of the held-out
PSNR (few-view training is chaotic) is known next to the difference of the means.

    python scripts/psnr_twin.py --iters 600 --seeds 0 1 2 --out gpurun_out/psnr_twin.json [--repo-modes split:split fp16:fp16]

Every run is a subprocess of oracle/twin.py; the scene is written once (seed 0) and shared by all runs."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import twin  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--seeds", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--res", type=int, default=400)
    ap.add_argument("--eval-views", type=int, default=4)
    ap.add_argument("--root", default="/tmp/cnerf_psnr_twin")
    ap.add_argument("--out", default="gpurun_out/psnr_twin.json")
    ap.add_argument("--repo-modes", nargs="+", default=["split:split", "fp16:fp16"], help="forward:grad precision pairs of the drop-in arm")
    ap.add_argument("--kind", default="blender")
    ap.add_argument("--train-views", type=int, default=3)
    a = ap.parse_args()
    twin.make_scene(a.kind, a.root, res=a.res, seed=0, n_train=a.train_views)
    runs = {"ref": []}
    t0 = time.time()
    for seed in a.seeds:
        r = twin.run_arm_subprocess("ref", a.kind, a.root, a.iters, seed=seed, eval_views=a.eval_views)
        runs["ref"].append({"seed": seed, **{k: r.get(k) for k in ("psnr", "psnr_views", "train_ms_per_iter", "error")}})
        print("ref", seed, r.get("psnr"), r.get("error"), flush=True)
    for mode in a.repo_modes:
        fp, gp = mode.split(":")
        runs[mode] = []
        for seed in a.seeds:
            r = twin.run_arm_subprocess("repo", a.kind, a.root, a.iters, seed=seed, eval_views=a.eval_views,
                                        env={"CNERF_FWD_PRECISION": fp, "CNERF_GRAD_PRECISION": gp})
            runs[mode].append({"seed": seed, **{k: r.get(k) for k in ("psnr", "psnr_views", "train_ms_per_iter", "error")}})
            print(mode, seed, r.get("psnr"), r.get("error"), flush=True)

    def stats(rows):
        v = [r["psnr"] for r in rows if r.get("psnr") is not None]
        return {"n": len(v), "mean": statistics.fmean(v) if v else None, "std": statistics.stdev(v) if len(v) > 1 else None,
                "min": min(v) if v else None, "max": max(v) if v else None,
                "train_ms_per_iter": statistics.fmean([r["train_ms_per_iter"] for r in rows if r.get("train_ms_per_iter")]) if v else None}
    summary = {"what": f"UNMODIFIED run_nerf.py train() x {a.iters} iters on one synthetic {a.kind} scene ({a.res}x{a.res}, {a.train_views} training views, N_rand 4096, 64 + 128 "
                       f"samples), held-out PSNR over {a.eval_views} views, seeds {a.seeds}; ref = PyTorch eager on the same GPU, the other arms = "
                       "consistentnerf_b200.dropin with forward:gradient precision as named",
               "arms": {k: stats(v) for k, v in runs.items()}}
    ref = summary["arms"]["ref"]
    for k, s in summary["arms"].items():
        if k != "ref" and s["mean"] is not None and ref["mean"] is not None:
            s["delta_db_vs_ref"] = s["mean"] - ref["mean"]
    summary["seconds"] = time.time() - t0
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump({"summary": summary, "runs": runs}, open(a.out, "w"), indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
