"""Generate the training-side fixtures by EXECUTING THE REFERENCE'S OWN CODE.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden_train.py

* ``adam.npz``: the optimizer the reference constructs at models/bin_model.py:97-100 is ``torch.optim.Adam`` (a
  third-party dependency, torch 2.11.0 here); three of its CPU steps on seeded tensors pin
  ``bin_oracle.adam_step`` and the CUDA kernel.
* ``blur_average.npz``: ``create_clips_overlap`` of data_scripts/adobe240fps/create_dataset_blur_N_frames_average.py
  is executed unmodified on a synthetic clip.  The script is not importable as is (module-level argparse, a trailing
  ``main()`` call, ``scipy.ndimage.imread`` / ``scipy.misc.imsave`` which no longer exist), so this driver parses the
  file, drops only the final ``main()`` statement, and supplies in-memory ``imread`` / ``imsave`` stand-ins; every
  arithmetic statement that runs is the reference's.  Nothing under /root/reference is copied into the repo.
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SCRIPT = "/root/reference/data_scripts/adobe240fps/create_dataset_blur_N_frames_average.py"


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrs.items()})


def adam_fixture():
    g = torch.Generator().manual_seed(21)
    shapes = [(3,), (12, 6, 3, 3), (4099,), (32, 8, 2, 4)]
    rec = {}
    for tag, wd in (("wd0", 0.0), ("wd", 1e-4)):
        ps = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in shapes]
        rec[f"{tag}_p0"] = np.concatenate([p.detach().numpy().ravel() for p in ps])
        opt = torch.optim.Adam(ps, lr=1e-4, weight_decay=wd, betas=(0.9, 0.99))      # yml lr_G / beta1 / beta2
        grads = []
        for _ in range(3):
            gs = [torch.randn(s, generator=g) * 0.01 for s in shapes]
            for p, gr in zip(ps, gs):
                p.grad = gr.clone()
            opt.step()
            grads.append(np.concatenate([gr.numpy().ravel() for gr in gs]))
        rec[f"{tag}_grads"] = np.stack(grads)
        rec[f"{tag}_p"] = np.concatenate([p.detach().numpy().ravel() for p in ps])
        rec[f"{tag}_m"] = np.concatenate([opt.state[p]["exp_avg"].numpy().ravel() for p in ps])
        rec[f"{tag}_v"] = np.concatenate([opt.state[p]["exp_avg_sq"].numpy().ravel() for p in ps])
    save("adam.npz", sizes=np.array([int(np.prod(s)) for s in shapes]), hyper=np.array([1e-4, 0.9, 0.99, 1e-8, 1e-4]), **rec)


def blur_fixture():
    T, H, W = 50, 6, 16
    rng = np.random.default_rng(22)
    frames = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    rec = {"frames": frames}
    tree = ast.parse(open(SCRIPT).read(), SCRIPT)
    last = tree.body[-1]
    assert isinstance(last, ast.Expr) and isinstance(last.value, ast.Call) and last.value.func.id == "main"
    tree.body = tree.body[:-1]
    code = compile(tree, SCRIPT, "exec")
    for ws in (7, 11):
        with tempfile.TemporaryDirectory() as tmp:
            root = os.path.join(tmp, "full_sharp")
            os.makedirs(os.path.join(root, "clip"))
            for t in range(T):
                open(os.path.join(root, "clip", f"{t + 1:05d}.png"), "wb").close()      # names only; pixels come from imread below
            saved = {}

            def imread(path):
                return frames[int(os.path.splitext(os.path.basename(path))[0]) - 1]

            def imsave(path, arr):
                saved[os.path.basename(path)] = np.array(arr)

            nd, misc = types.ModuleType("scipy.ndimage"), types.ModuleType("scipy.misc")
            nd.imread, misc.imsave = imread, imsave
            old = {k: sys.modules.get(k) for k in ("scipy.ndimage", "scipy.misc")}
            argv = sys.argv
            sys.modules["scipy.ndimage"], sys.modules["scipy.misc"] = nd, misc
            sys.argv = ["x", "--ffmpeg_dir", tmp, "--videos_folder", tmp, "--dataset_folder", tmp, "--window_size", str(ws)]
            try:
                ns = {"__name__": "reference_blur_script"}
                exec(code, ns)
                os.makedirs(os.path.join(tmp, "lists"))
                ns["create_clips_overlap"](["clip.mp4"], root, os.path.join(tmp, "train"), os.path.join(tmp, "train_blur"),
                                           os.path.join(tmp, "lists"))
            finally:
                sys.argv = argv
                for k, v in old.items():
                    if v is None:
                        sys.modules.pop(k, None)
                    else:
                        sys.modules[k] = v
            names = sorted(saved)
            rec[f"ws{ws}_mid"] = np.array([int(os.path.splitext(n)[0]) - 1 for n in names])       # 0-based centre frame
            rec[f"ws{ws}_out"] = np.stack([saved[n] for n in names])
            assert rec[f"ws{ws}_out"].dtype == np.uint8
    save("blur_average.npz", **rec)


if __name__ == "__main__":
    adam_fixture()
    blur_fixture()
