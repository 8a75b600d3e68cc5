"""The caller side of the render path: one optimisation step and density control.

SURVEY 8f ranks 1-2 ("next" after the hot path).  The reference documents this loop but does not
implement it: `GaussianTrainer.train_step` is `pass` (src/train/trainer.py:61-65) and
`DensityController.densify_and_prune` raises (src/core/optimizer.py:64 calls a property,
:70 never clears the state, :71 calls a missing method).  What is restated here is the working
intent: five Adam parameter groups (optimizer.py:100-109), cosine learning-rate decay with a
delayed start for positions (optimizer.py:21-32), densify every `densify_interval` iterations
between `densify_from_iter` and `densify_until_iter` (optimizer.py:39-41), split / clone / prune
opacity <= 0.01 (optimizer.py:61-66) and a rebuilt optimiser afterwards (optimizer.py:137).
Everything here is host-side torch on device tensors; the kernels are the renderer's.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch


@dataclass
class TrainingConfig:
    """The reference's dataclass, field for field (config/config.py:33-67), so `TrainingConfig(**yaml)` of a reference
    config file constructs; `learning_rate` and `batch_size` are read by nothing there either."""
    data_path: str = "data/scene"
    output_path: str = "output"
    iterations: int = 30000
    learning_rate: float = 0.0025
    batch_size: int = 1
    position_lr_init: float = 0.00016
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 30000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    densify_from_iter: int = 500
    densify_until_iter: int = 15000
    densify_grad_threshold: float = 0.0002
    densify_interval: int = 100
    image_height: int = 800
    image_width: int = 800
    device: str = "cuda"


class ConfigManager:
    """config/config.py:69-94: the training configuration as a YAML file."""

    @staticmethod
    def load_from_yaml(config_path: str) -> TrainingConfig:
        import yaml
        with open(config_path, "r", encoding="utf-8") as f:
            return TrainingConfig(**(yaml.safe_load(f) or {}))

    @staticmethod
    def save_to_yaml(config: TrainingConfig, config_path: str) -> None:
        import yaml
        os.makedirs(os.path.dirname(os.path.abspath(config_path)), exist_ok=True)
        with open(config_path, "w", encoding="utf-8") as f:
            yaml.safe_dump({k: getattr(config, k) for k in config.__dataclass_fields__}, f, allow_unicode=True)

    @staticmethod
    def get_default_config() -> TrainingConfig:
        return TrainingConfig()


class LearningRateScheduler:
    """Cosine decay from lr_init to lr_final over max_steps, scaled during the first
    lr_delay_steps by a ramp from lr_delay_mult to 1 (optimizer.py:7-32)."""

    def __init__(self, lr_init: float, lr_final: float, lr_delay_steps: int, lr_delay_mult: float, max_steps: int):
        self.lr_init, self.lr_final = lr_init, lr_final
        self.lr_delay_steps, self.lr_delay_mult, self.max_steps = lr_delay_steps, lr_delay_mult, max_steps

    def get_lr(self, step: int) -> float:
        if self.max_steps <= 0:
            return self.lr_final
        t = min(step, self.max_steps) / self.max_steps
        lr = self.lr_final + (self.lr_init - self.lr_final) * 0.5 * (1.0 + math.cos(math.pi * t))
        if self.lr_delay_steps > 0:
            lr *= self.lr_delay_mult + (1.0 - self.lr_delay_mult) * min(step / self.lr_delay_steps, 1.0)
        return float(lr)


class GaussianOptimizer:
    """Adam over the five trained parameter groups (the SH rest block is left out while colour is
    DC-only: its gradient is identically zero)."""

    GROUPS = (("xyz", "_xyz", "position_lr_init"), ("f_dc", "_features_dc", "feature_lr"),
              ("opacity", "_opacity", "opacity_lr"), ("scaling", "_scaling", "scaling_lr"),
              ("rotation", "_rotation", "rotation_lr"))

    def __init__(self, model, config: Optional[TrainingConfig] = None):
        self.model, self.config = model, config or TrainingConfig()
        self.xyz_scheduler = LearningRateScheduler(self.config.position_lr_init, self.config.position_lr_final,
                                                   int(0.01 * self.config.position_lr_max_steps),
                                                   self.config.position_lr_delay_mult, self.config.position_lr_max_steps)
        self.rebuild()

    def rebuild(self) -> None:
        """Fresh Adam on the model's current parameters; state is discarded, as in the reference."""
        groups = [{"params": [getattr(self.model, attr)], "lr": getattr(self.config, lr), "name": name}
                  for name, attr, lr in self.GROUPS]
        # one fused multi-tensor kernel per step on CUDA parameters; plain Adam on CPU (tests)
        fused = all(g["params"][0].is_cuda for g in groups)
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15, fused=fused)

    @torch.no_grad()
    def carry_state(self, src_row: torch.Tensor) -> None:
        """Rebuild Adam on the model's new parameters after a densification round and CARRY the moments over:
        row j of every new exp_avg / exp_avg_sq is the old row `src_row[j]`, or zero where `src_row[j] < 0` (clone
        copies and split children start without history, as appended rows do in the original 3DGS trainer); the step
        counts and learning rates are kept.  The reference's intent at optimizer.py:67-71 is the cruder `rebuild()`
        (state cleared); SURVEY 8f rank 1 asks for the carry-over."""
        old = self.optimizer
        saved = {g["name"]: (g["lr"], dict(old.state.get(g["params"][0], {}))) for g in old.param_groups}
        self.rebuild()
        src_row = src_row.to(torch.int64)
        dst = torch.nonzero(src_row >= 0, as_tuple=False).squeeze(-1)
        src = src_row[dst]
        for g in self.optimizer.param_groups:
            lr, st = saved[g["name"]]
            g["lr"] = lr
            if not st:
                continue
            p = g["params"][0]
            if p.shape[0] != src_row.shape[0]:
                raise ValueError(f"src_row has {src_row.shape[0]} rows, parameter {g['name']} has {p.shape[0]}")
            new = {}
            for k, v in st.items():
                if torch.is_tensor(v) and v.dim() > 0 and v.shape[1:] == p.shape[1:]:
                    t = torch.zeros_like(p)
                    t[dst] = v.to(p.device)[src.to(v.device)]
                    new[k] = t
                else:
                    new[k] = v.clone() if torch.is_tensor(v) else v       # the step counter
            self.optimizer.state[p] = new

    def update_learning_rate(self, iteration: int) -> float:
        lr = self.xyz_scheduler.get_lr(iteration)
        for g in self.optimizer.param_groups:
            if g["name"] == "xyz":
                g["lr"] = lr
        return lr

    def step(self) -> None:
        self.optimizer.step()

    def zero_grad(self) -> None:
        self.optimizer.zero_grad(set_to_none=True)


class DensityController:
    def __init__(self, config: Optional[TrainingConfig] = None, fused: bool = True, carry_state: bool = True):
        self.config = config or TrainingConfig()
        self.fused = fused           # CUDA models: gs_densify_plan / gs_densify_apply instead of tensor ops
        self.carry_state = carry_state   # False: the optimiser forgets its moments at every round (optimizer.py:67-71)

    def _refresh_optimizer(self, optimizer, src_row) -> None:
        if optimizer is None:
            return
        if self.carry_state:
            optimizer.carry_state(src_row)
        else:
            optimizer.rebuild()

    def should_densify(self, iteration: int) -> bool:
        c = self.config
        return c.densify_from_iter <= iteration <= c.densify_until_iter and iteration % c.densify_interval == 0

    @torch.no_grad()
    def densify_and_prune(self, model, optimizer: Optional[GaussianOptimizer], scene_extent: float,
                          grad: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> Dict[str, int]:
        """Split, clone, prune; `grad` defaults to `_xyz.grad` as in the reference.  Both masks are
        taken from the same pre-densification gradient (the split re-creates the parameters, so the
        reference's second read of `.grad` would find none)."""
        g = model._xyz.grad if grad is None else grad
        if g is None:
            return {"split": 0, "cloned": 0, "pruned": 0, "points": model.get_num_points()}
        th = self.config.densify_grad_threshold
        if g.is_cuda and self.fused and hasattr(model, "densify_fused"):
            # one plan + one apply pass on the device; the counts below are those of the sequential formulation
            # except "pruned", which here counts only originals (a pruned clone/child was never created)
            n0 = model.get_num_points()
            r = model.densify_fused(g, th, scene_extent, 0.01, generator=generator)
            self._refresh_optimizer(optimizer, r["src_row"])
            return {"split": r["split"], "cloned": r["cloned"], "pruned": n0 - r["kept"] - r["split"], "points": r["points"],
                    "src_row": r["src_row"]}
        gn = g.norm(dim=-1)
        sig = model.get_scaling.mean(dim=-1)
        clone_mask = (gn > th) & (sig < 0.01 * scene_extent)
        split_mask = (gn > th) & (sig > 0.03 * scene_extent)
        # clone first: appended rows do not disturb the indices the split mask refers to
        n0 = model.get_num_points()
        # where every row of the new model comes from: the original's index, or -1 for rows created here
        src_row = torch.arange(n0, device=g.device, dtype=torch.int64)
        fresh = lambda k: torch.full((k,), -1, device=g.device, dtype=torch.int64)  # noqa: E731
        cloned = model.density_and_clone(th, scene_extent, grad=g, generator=generator)
        if cloned:
            split_mask = torch.cat([split_mask, torch.zeros(cloned, dtype=torch.bool, device=split_mask.device)])
            g = torch.cat([g, torch.zeros(cloned, 3, device=g.device, dtype=g.dtype)])
            src_row = torch.cat([src_row, fresh(cloned)])
        split = 0
        if bool(split_mask.any()):
            g_for_split = torch.where(split_mask.unsqueeze(-1), g, torch.zeros_like(g))
            split = model.density_and_split(th, scene_extent, grad=g_for_split)
            src_row = torch.cat([src_row[~split_mask], fresh(2 * split)])
        keep = model.get_opacity.squeeze(1) > 0.01
        pruned = int((~keep).sum())
        if pruned:
            model.prune_points(keep)
            src_row = src_row[keep]
        self._refresh_optimizer(optimizer, src_row)
        assert int(clone_mask.sum()) == cloned and n0 + cloned + split - pruned == model.get_num_points() == src_row.shape[0]
        return {"split": split, "cloned": cloned, "pruned": pruned, "points": model.get_num_points(), "src_row": src_row}


def l1_loss(image: torch.Tensor, target: torch.Tensor):
    """The L1 term of the reference's loss (loss.py:52); its SSIM term does not run (loss.py:26,39).
    On the device: one fused kernel that also writes the gradient plane (losses.l1_loss, csrc/loss.cu), so the
    loss's backward launches nothing.  Host tensors (the CPU tests of the step logic with a stand-in renderer) go
    through the two torch ops of the reference."""
    if image.is_cuda:
        from .losses import l1_loss as fused_l1
        return fused_l1(image, target)
    return (image - target).abs().mean()


def train_step(model, renderer, camera, target: torch.Tensor, optimizer: GaussianOptimizer, settings, iteration: int,
               loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor] = l1_loss) -> Dict[str, object]:
    """sample -> render -> loss -> backward -> Adam step (the intent of trainer.py:45-65), plus the
    densification statistics the reference allocates but never fills (gaussian_model.py:29-31)."""
    optimizer.update_learning_rate(iteration)
    optimizer.zero_grad()
    out = renderer.render(camera, model, settings)
    out["viewspace_points"].retain_grad()
    loss = loss_fn(out["image"], target)
    loss.backward()
    with torch.no_grad():
        model.add_densification_stats(out["viewspace_points"].grad, out["visibility_filter"], out["radii"])
    optimizer.step()
    return {"loss": loss.detach(), "out": out}


def densification_stress(model, renderer, cameras, settings, controller: DensityController, scene_extent: float,
                         target_points: int, max_rounds: int = 64, grad_scale: float = 1.0) -> Dict[str, object]:
    """BASELINE config[4]: grow the model from its current size to >= target_points by repeated
    render -> backward -> statistics -> split/clone/prune rounds (no optimiser step: the parameters
    only change through densification, so the run is reproducible)."""
    history = []
    for r in range(max_rounds):
        if model.get_num_points() >= target_points:
            break
        for p in (model._xyz, model._scaling, model._rotation, model._opacity, model._features_dc, model._features_rest):
            p.grad = None
        cam = cameras[r % len(cameras)]
        out = renderer.render(cam, model, settings)
        out["viewspace_points"].retain_grad()
        (out["image"].mean() * grad_scale).backward()
        with torch.no_grad():
            model.add_densification_stats(out["viewspace_points"].grad, out["visibility_filter"], out["radii"])
        stats = controller.densify_and_prune(model, None, scene_extent)
        stats["round"] = r
        stats["tile_pairs"] = renderer.last_stats.get("tile_pairs", 0)
        history.append(stats)
        if stats["split"] + stats["cloned"] == 0:
            break
    return {"points": model.get_num_points(), "history": history}


class GaussianTrainer:
    """The reference's trainer surface -- `setup / train / train_step / validate / save_checkpoint / load_checkpoint /
    get_scene_extent` (src/train/trainer.py:12-97) -- with the loop its docstrings describe; in the reference every one
    of these methods is `pass`.  One iteration = pick a training camera, render it, L1 against the camera's image,
    backward, Adam step, densification statistics, and density control on its schedule (trainer.py:45-59).

    The reference's dataset classes are out of scope: cameras are passed in, each carrying its ground-truth
    image [3,H,W] as `camera._image` (what `Camera(..., image=...)` stores).  Any renderer with the reference's
    `render(camera, gaussians, settings)` contract works; the default is the B200 one."""

    def __init__(self, config: Optional[TrainingConfig] = None, cameras: Optional[Sequence] = None, gaussians=None,
                 renderer=None, val_cameras: Optional[Sequence] = None, bg_color: Sequence[float] = (0.0, 0.0, 0.0), seed: int = 0):
        self.config = config or TrainingConfig()
        self.cameras = list(cameras or [])
        self.val_cameras = list(val_cameras or [])
        self.gaussians = gaussians
        self.renderer = renderer
        self.optimizer: Optional[GaussianOptimizer] = None
        self.density_controller: Optional[DensityController] = None
        self.loss_fn = l1_loss
        self.bg_color = tuple(float(v) for v in bg_color)
        self.iteration = 0
        self.scene_extent = 0.0
        self.train_losses: List[float] = []
        self.val_losses: List[float] = []
        self._order = torch.Generator().manual_seed(seed)          # camera sampling: reproducible, same on every rank

    # ---- trainer.py:31-43 ---------------------------------------------------------------------------
    def setup(self) -> None:
        """Model (the one passed in, else from `<data_path>/points3D.*` when present, else `create_from_random`),
        renderer, optimiser, density controller, scene extent."""
        from .renderer import GaussianRenderer
        from .scene import GaussianModel
        if self.gaussians is None:
            self.gaussians = GaussianModel(self.config)
            pcd = next((os.path.join(self.config.data_path, f) for f in ("points3D.ply", "points3D.txt", "points3D.npz")
                        if os.path.exists(os.path.join(self.config.data_path, f))), None)
            if pcd is not None:
                self.gaussians.create_from_pcd(pcd)
            else:
                self.gaussians.create_from_random(10000, 1.0)
        if self.renderer is None:
            self.renderer = GaussianRenderer()
        self.optimizer = GaussianOptimizer(self.gaussians, self.config)
        self.density_controller = DensityController(self.config)
        self.scene_extent = self.get_scene_extent()

    def get_scene_extent(self) -> float:
        """Radius of the camera rig, the scale the densification thresholds are relative to (optimizer.py:61-63 compare
        sigma with 0.01 / 0.03 x this): 1.1 x the largest distance of a camera centre from the centroid of the centres;
        with fewer than two cameras, half the diagonal of the model's bounding box."""
        centres = [c.camera_center for c in self.cameras if hasattr(c, "camera_center")]
        if len(centres) >= 2:
            C = torch.stack([torch.as_tensor(c, dtype=torch.float32).reshape(3).cpu() for c in centres])
            return 1.1 * float((C - C.mean(dim=0)).norm(dim=-1).max())
        if self.gaussians is not None and self.gaussians.get_num_points() > 0:
            xyz = self.gaussians.get_xyz.detach()
            return 0.5 * float((xyz.max(dim=0).values - xyz.min(dim=0).values).norm())
        return 1.0

    # ---- trainer.py:45-65 ---------------------------------------------------------------------------
    def _settings(self, camera):
        from .renderer import RenderSettings
        dev = self.gaussians.get_xyz.device
        return RenderSettings(int(camera._height), int(camera._width), torch.tensor(self.bg_color, dtype=torch.float32, device=dev))

    def _target(self, camera) -> torch.Tensor:
        img = getattr(camera, "_image", None)
        if img is None:
            raise ValueError("training cameras carry their ground-truth image as camera._image ([3,H,W])")
        return img.to(self.gaussians.get_xyz.device, torch.float32)

    def train_step(self, camera) -> Dict[str, float]:
        """One iteration on `camera` (trainer.py:61-65); returns floats, as the reference's signature says (one host
        read of the loss per step -- `training.train_step` is the variant that leaves the loss on the device)."""
        self.iteration += 1
        res = train_step(self.gaussians, self.renderer, camera, self._target(camera), self.optimizer, self._settings(camera),
                         self.iteration, self.loss_fn)
        stats: Dict[str, float] = {"iteration": float(self.iteration), "points": float(self.gaussians.get_num_points())}
        if self.density_controller.should_densify(self.iteration):
            n = self.gaussians
            grad = n.xyz_gradient_accum / n.denom.clamp_min(1.0)           # mean |dL/d means2D| over the views that saw the splat
            d = self.density_controller.densify_and_prune(n, self.optimizer, self.scene_extent,
                                                          grad=torch.cat([grad[:, :1], torch.zeros_like(grad[:, 1:])], dim=1))
            stats.update({k: float(d[k]) for k in ("split", "cloned", "pruned")})
            stats["points"] = float(n.get_num_points())
        stats["loss"] = float(res["loss"])
        return stats

    def train(self, iterations: Optional[int] = None, val_every: int = 0, checkpoint_every: int = 0) -> None:
        """Main loop (trainer.py:45-59): random training camera, step, density control, periodic validation / checkpoint."""
        if self.optimizer is None:
            self.setup()
        if not self.cameras:
            raise ValueError("no training cameras")
        total = self.config.iterations if iterations is None else int(iterations)
        while self.iteration < total:
            cam = self.cameras[int(torch.randint(len(self.cameras), (1,), generator=self._order))]
            self.train_losses.append(self.train_step(cam)["loss"])
            if val_every and self.iteration % val_every == 0:
                self.val_losses.append(self.validate()["l1"])
            if checkpoint_every and self.iteration % checkpoint_every == 0:
                self.save_checkpoint(self.iteration)

    @torch.no_grad()
    def validate(self) -> Dict[str, float]:
        """Mean L1 and PSNR over the validation cameras (the training cameras when none were given)."""
        cams = self.val_cameras or self.cameras
        l1s, psnrs = [], []
        for cam in cams:
            img = self.renderer.render(cam, self.gaussians, self._settings(cam))["image"]
            tgt = self._target(cam)
            l1s.append(float((img - tgt).abs().mean()))
            psnrs.append(float(-10.0 * torch.log10(((img - tgt) ** 2).mean().clamp_min(1e-12))))
        return {"l1": sum(l1s) / len(l1s), "psnr": sum(psnrs) / len(psnrs), "views": float(len(cams))}

    # ---- trainer.py:73-83 ---------------------------------------------------------------------------
    def _checkpoint_path(self, iteration: int) -> str:
        return os.path.join(self.config.output_path, f"checkpoint_{int(iteration):07d}.pt")

    def save_checkpoint(self, iteration: int) -> None:
        """Parameters, Adam moments and step counts, the densification statistics, the iteration and the camera-sampling
        state: `load_checkpoint` followed by `train()` continues exactly where this run would have."""
        os.makedirs(self.config.output_path, exist_ok=True)
        g = self.gaussians
        torch.save({
            "iteration": int(self.iteration), "scene_extent": float(self.scene_extent),
            "params": {name: getattr(g, name).detach().cpu() for name in
                       ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")},
            "stats": {"xyz_gradient_accum": g.xyz_gradient_accum.cpu(), "denom": g.denom.cpu(), "max_radii2D": g.max_radii2D.cpu()},
            "optimizer": self.optimizer.optimizer.state_dict(), "order": self._order.get_state(),
            "train_losses": list(self.train_losses), "val_losses": list(self.val_losses),
        }, self._checkpoint_path(iteration))

    def load_checkpoint(self, iteration: int) -> None:
        ck = torch.load(self._checkpoint_path(iteration), map_location="cpu", weights_only=False)
        if self.gaussians is None:
            from .scene import GaussianModel
            self.gaussians = GaussianModel(self.config)
        p = ck["params"]
        self.gaussians.create_from_tensors(p["_xyz"], p["_features_dc"], p["_scaling"], p["_rotation"], p["_opacity"], p["_features_rest"])
        dev = self.gaussians.get_xyz.device
        for k, v in ck["stats"].items():
            setattr(self.gaussians, k, v.to(dev))
        if self.renderer is None:
            from .renderer import GaussianRenderer
            self.renderer = GaussianRenderer()
        self.optimizer = GaussianOptimizer(self.gaussians, self.config)
        self.optimizer.optimizer.load_state_dict(ck["optimizer"])
        self.density_controller = DensityController(self.config)
        self.iteration, self.scene_extent = int(ck["iteration"]), float(ck["scene_extent"])
        self._order.set_state(ck["order"])
        self.train_losses, self.val_losses = list(ck["train_losses"]), list(ck["val_losses"])
