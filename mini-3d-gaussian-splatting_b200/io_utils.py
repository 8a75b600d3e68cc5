"""Scene I/O and camera helpers around the render path.

Mirrors the working parts of the reference's `src/utils/io_utils.py` (`IOUtils.save_image` :17-23,
`IOUtils.load_point_cloud` :33-85) and `src/core/camera.py` (`CameraUtils.build_world_view_matrix` :79-141) --
same names, arguments and file formats -- and fills the stubs next to them (`load_image`, `save_point_cloud`)
plus what a trained model needs: `save_gaussians` / `load_gaussians` (npz of the six parameter tensors) and
an ASCII/binary-little-endian PLY reader/writer for point clouds.  Host-side numpy/torch only."""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

try:  # Pillow is optional, exactly as in the reference (io_utils.py:7-12)
    from PIL import Image
except ImportError:  # pragma: no cover
    Image = None


class CameraUtils:
    @staticmethod
    def build_world_view_matrix(R_np: np.ndarray, T_np: np.ndarray, from_c2w: bool, device=None, dtype=None) -> torch.Tensor:
        """4x4 world-to-camera matrix with X_c = R_wc X_w + t_wc (camera.py:79-141).
        from_c2w=True: inputs are the camera-to-world rotation R_cw and the camera centre C_w
        (R_wc = R_cw^T, t_wc = -R_cw^T C_w); False: inputs are R_wc and t_wc themselves."""
        R_np = np.asarray(R_np)
        assert R_np.shape == (3, 3), f"R should be [3,3], got {R_np.shape}"
        R = torch.from_numpy(R_np.copy())
        T = torch.from_numpy(np.asarray(T_np).reshape(3, 1).copy())
        if dtype is not None:
            R, T = R.to(dtype), T.to(dtype)
        if device is not None:
            R, T = R.to(device), T.to(device)
        view = torch.eye(4, device=R.device, dtype=R.dtype)
        if from_c2w:
            R_wc = R.transpose(0, 1)
            t_wc = -(R_wc @ T).flatten()
        else:
            R_wc, t_wc = R, T.flatten()
        view[:3, :3] = R_wc
        view[:3, 3] = t_wc
        return view

    @staticmethod
    def build_projection_matrix(znear: float, zfar: float, fovX: float, fovY: float,
                                width: Optional[int] = None, height: Optional[int] = None) -> torch.Tensor:
        """OpenGL-style perspective matrix of camera.py:143-188 (the reference defines it twice under one name -- once from
        the half-angle tangents, once from focal lengths and the image size -- and both give the same matrix, so `width` /
        `height` are accepted and not needed).  The renderer never reads it (renderer.py:140-152 takes fx, fy from the
        fields of view); it is here for callers that project points themselves."""
        import math
        tx = max(abs(math.tan(fovX * 0.5)), 1e-6)
        ty = max(abs(math.tan(fovY * 0.5)), 1e-6)
        P = torch.zeros(4, 4, dtype=torch.float32)
        P[0, 0] = 1.0 / tx
        P[1, 1] = 1.0 / ty
        P[2, 2] = -(zfar + znear) / (zfar - znear)
        P[2, 3] = -(2.0 * zfar * znear) / (zfar - znear)
        P[3, 2] = -1.0
        return P

    @staticmethod
    def focal_to_fov(focal: float, pixels: int) -> float:
        return 2.0 * float(np.arctan(pixels / (2.0 * focal)))

    @staticmethod
    def fov_to_focal(fov: float, pixels: int) -> float:
        return pixels / (2.0 * float(np.tan(fov / 2.0)))


class IOUtils:
    # ---- images ---------------------------------------------------------------------------------
    @staticmethod
    def save_image(image: torch.Tensor, path: str) -> None:
        """[3,H,W] float image in [0,1] -> 8-bit file (io_utils.py:17-23); .npy when Pillow is absent."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        arr = (image.detach().cpu().clamp(0, 1) * 255).byte().permute(1, 2, 0).numpy()
        if Image is None or str(path).endswith(".npy"):
            np.save(str(path) if str(path).endswith(".npy") else str(path) + ".npy", arr)
            return
        Image.fromarray(arr).save(path)

    @staticmethod
    def load_image(path: str) -> torch.Tensor:
        """File -> [3,H,W] float32 in [0,1] (the reference leaves this a stub, io_utils.py:25-27)."""
        if str(path).endswith(".npy"):
            arr = np.load(str(path))
        else:
            if Image is None:
                raise RuntimeError("Pillow is not installed")
            arr = np.asarray(Image.open(path).convert("RGB"))
        return torch.from_numpy(arr.astype(np.float32) / 255.0).permute(2, 0, 1).contiguous()

    # ---- point clouds ---------------------------------------------------------------------------
    @staticmethod
    def load_point_cloud(path: str) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """points [N,3] float32 and colours [N,3] float32 in [0,1] or None.  Formats of the reference
        (io_utils.py:33-85): .npz (`points`, `colors`), .npy ([N,3] or [N,>=6]), COLMAP points3D.txt, plain
        `x y z [r g b]` text; plus .ply (ascii or binary_little_endian, uchar or float colours)."""
        p = Path(path)
        suf = p.suffix.lower()
        if suf == ".npz":
            data = np.load(str(p))
            pts = data["points"] if "points" in data else np.zeros((0, 3), np.float32)
            cols = data["colors"] if "colors" in data else None
            return pts.astype(np.float32), (None if cols is None else cols.astype(np.float32))
        if suf == ".npy":
            arr = np.load(str(p))
            if arr.ndim == 2 and arr.shape[1] >= 6:
                return arr[:, :3].astype(np.float32), arr[:, 3:6].astype(np.float32)
            return arr[:, :3].astype(np.float32), None
        if suf == ".ply":
            return _read_ply(p)
        colmap = suf == ".txt" and p.name == "points3D.txt"
        pts, cols = [], []
        with open(p, "r", encoding="utf-8", errors="ignore") as f:
            for line in f:
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                parts = line.split()
                try:
                    if colmap:                      # POINT3D_ID X Y Z R G B ERROR TRACK[]
                        if len(parts) < 8:
                            continue
                        pts.append([float(v) for v in parts[1:4]])
                        cols.append([float(v) / 255.0 for v in parts[4:7]])
                    else:
                        vals = [float(v) for v in parts]
                        if len(vals) >= 3:
                            pts.append(vals[:3])
                            if len(vals) >= 6:
                                cols.append(vals[3:6])
                except ValueError:
                    continue
        pts_a = np.asarray(pts, dtype=np.float32).reshape(-1, 3)
        cols_a = np.asarray(cols, dtype=np.float32).reshape(-1, 3) if len(cols) == len(pts) and cols else None
        return pts_a, cols_a

    load_pcd = load_point_cloud                     # the name gaussian_model.py:44 calls

    @staticmethod
    def save_point_cloud(points: np.ndarray, colors: Optional[np.ndarray], path: str) -> None:
        """.npz, .npy or binary .ply (the reference leaves this a stub, io_utils.py:29-31)."""
        p = Path(path)
        p.parent.mkdir(parents=True, exist_ok=True)
        pts = np.asarray(points, dtype=np.float32).reshape(-1, 3)
        cols = None if colors is None else np.asarray(colors, dtype=np.float32).reshape(-1, 3)
        suf = p.suffix.lower()
        if suf == ".npz":
            np.savez(str(p), points=pts, **({} if cols is None else {"colors": cols}))
        elif suf == ".npy":
            np.save(str(p), pts if cols is None else np.concatenate([pts, cols], axis=1))
        elif suf == ".ply":
            _write_ply(p, pts, cols)
        else:
            np.savetxt(str(p), pts if cols is None else np.concatenate([pts, cols], axis=1), fmt="%.7g")

    # ---- trained models ---------------------------------------------------------------------------
    PARAMS = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")

    @staticmethod
    def save_gaussians(model, path: str) -> None:
        """The six parameter tensors of a GaussianModel as one .npz (raw values: log-scales, opacity logits)."""
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        np.savez(str(path), **{k: getattr(model, k).detach().cpu().numpy() for k in IOUtils.PARAMS})

    @staticmethod
    def load_gaussians(model, path: str):
        data = np.load(str(path))
        dev = model._xyz.device
        t = {k: torch.from_numpy(data[k]).to(dev) for k in IOUtils.PARAMS}
        model.create_from_tensors(t["_xyz"], t["_features_dc"], t["_scaling"], t["_rotation"], t["_opacity"], t["_features_rest"])
        return model


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def _read_ply(p: Path):
    with open(p, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{p}: not a PLY file")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{p}: truncated header")
            tok = line.decode("ascii", "ignore").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{p}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(props)))
            cols = {name: rows[:, i] for i, (name, _) in enumerate(props)}
        elif fmt == "binary_little_endian":
            dt = np.dtype([(name, "<" + t) for name, t in props])
            rec = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            cols = {name: rec[name] for name, _ in props}
        else:
            raise ValueError(f"{p}: PLY format {fmt} not supported")
    pts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1).astype(np.float32)
    colour = None
    if all(k in cols for k in ("red", "green", "blue")):
        colour = np.stack([cols["red"], cols["green"], cols["blue"]], axis=1).astype(np.float32)
        if dict(props)["red"] == "u1" or colour.max(initial=0.0) > 1.0:
            colour = colour / 255.0
    return pts, colour


def _write_ply(p: Path, pts: np.ndarray, cols: Optional[np.ndarray]) -> None:
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if cols is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.zeros(len(pts), dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if cols is not None:
        c8 = np.clip(np.rint(cols * 255.0), 0, 255).astype(np.uint8)
        rec["red"], rec["green"], rec["blue"] = c8[:, 0], c8[:, 1], c8[:, 2]
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(pts)}", "property float x", "property float y",
              "property float z"]
    if cols is not None:
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    header.append("end_header")
    with open(p, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())
