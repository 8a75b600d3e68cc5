#!/usr/bin/env python
"""BASELINE configs[0] as a script (the reference ships this file empty): 10 000 random-init Gaussians, 256x256, one
synthetic camera, render() forward + backward -- here on the B200 renderer, same calls as with the reference's.

    python examples/simple_scene.py [--splats 10000] [--size 256] [--out simple_scene.png]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gsplat_b200 as gb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--splats", type=int, default=10000)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--out", default="simple_scene.png")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        sys.exit("simple_scene.py needs a CUDA device: the renderer has no CPU path")
    model = gb.GaussianModel(device="cuda")
    model.create_from_random(args.splats, 1.0, seed=0)
    camera = gb.Camera.look_at_origin_c0(args.size, args.size)
    settings = gb.RenderSettings(image_height=args.size, image_width=args.size, bg_color=torch.zeros(3, device="cuda"))
    renderer = gb.GaussianRenderer()
    for it in range(3):                                     # the first frame sizes the binning buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = renderer.render(camera, model, settings)
        out["viewspace_points"].retain_grad()
        out["image"].mean().backward()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    print(f"{args.splats} splats, {args.size}x{args.size}: forward + backward {ms:.2f} ms; "
          f"visible {int(out['visibility_filter'].sum())}, max |dL/d means2D| {float(out['viewspace_points'].grad.abs().max()):.3e}")
    gb.IOUtils.save_image(out["image"].detach().cpu(), args.out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
