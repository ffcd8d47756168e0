"""Procedural "Lego-shaped" NeRF scene (BASELINE.json configs[1]).

A brick-like union of axis-aligned boxes and studs inside the unit cube, ray-cast analytically,
seen from cameras on the upper hemisphere at radius 4.0 (NeRF convention, as nerf_synthetic) with
camera_angle_x = 0.6911, RGBA8 output with alpha. The loader transform is the reference's
nerf_matrix_to_ngp (include/neural-graphics-primitives/nerf_loader.h:113-132) with the
nerf_synthetic values scale = 0.33, offset = 0.5 that the reference's docs name
(docs/nerf_dataset_tips.md:22). Everything is deterministic.

The ray definition equals the training kernel's (src/testbed_nerf.cu:1174-1189): pixel (u,v) in [0,1)^2,
d_cam = ((u-cx)*w/fx, (v-cy)*h/fy, 1), d_world = R d_cam, origin = t.
"""
import json
import math
import os

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618
NERF_SCALE = 0.33
NERF_OFFSET = 0.5


def nerf_matrix_to_ngp(c2w, scale=NERF_SCALE, offset=NERF_OFFSET):
    """nerf_loader.h:113-132 (from_mitsuba = false, scale_columns = false). c2w: [3,4] or [4,4]."""
    m = np.array(c2w, dtype=np.float32)[:3, :4].copy()
    m[:, 1] *= -1
    m[:, 2] *= -1
    m[:, 3] = m[:, 3] * np.float32(scale) + np.float32(offset)
    return m[[1, 2, 0], :].copy()  # cycle axes xyz <- yzx


def hemisphere_cameras(n, radius=4.0, seed=0):
    """n camera-to-world matrices (NeRF/OpenGL convention: x right, y up, z back) looking at the origin."""
    out = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        z = 0.15 + 0.8 * (i + 0.5) / n  # upper hemisphere
        r = math.sqrt(max(0.0, 1.0 - z * z))
        th = golden * i + 0.1 * seed
        pos = np.array([r * math.cos(th), r * math.sin(th), z], dtype=np.float64) * radius
        back = pos / np.linalg.norm(pos)
        up_w = np.array([0.0, 0.0, 1.0])
        right = np.cross(up_w, back); right /= np.linalg.norm(right)
        up = np.cross(back, right)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, pos
        out.append(m.astype(np.float32))
    return out


def lego_boxes():
    """(min, max, rgb) boxes in the ngp unit cube: base plate, two bricks, an arm, studs."""
    boxes = []
    def add(cx, cy, cz, sx, sy, sz, col):
        boxes.append(((cx - sx / 2, cy - sy / 2, cz - sz / 2), (cx + sx / 2, cy + sy / 2, cz + sz / 2), col))
    # NOTE ngp axes: the loader cycles xyz <- yzx, "up" (NeRF z) becomes ngp y.
    add(0.50, 0.30, 0.50, 0.46, 0.04, 0.34, (0.25, 0.55, 0.25))   # base plate
    add(0.44, 0.38, 0.50, 0.26, 0.12, 0.18, (0.85, 0.70, 0.10))   # yellow brick
    add(0.58, 0.47, 0.46, 0.14, 0.06, 0.12, (0.80, 0.15, 0.12))   # red brick
    add(0.62, 0.40, 0.58, 0.06, 0.16, 0.06, (0.55, 0.55, 0.60))   # pillar
    add(0.50, 0.56, 0.58, 0.30, 0.04, 0.05, (0.20, 0.30, 0.80))   # arm
    for ix in range(4):
        for iz in range(3):
            add(0.35 + 0.06 * ix, 0.45, 0.45 + 0.05 * iz, 0.03, 0.02, 0.03, (0.90, 0.78, 0.15))  # studs
    return boxes


def _linear_to_srgb(x):
    return torch.where(x < 0.0031308, 12.92 * x, 1.055 * torch.clamp(x, min=1e-8) ** (1.0 / 2.4) - 0.055)


def render_image(ngp_xform, res, fx, fy, boxes, device="cpu"):
    """Analytic ray-cast of the box scene -> uint8 [res,res,4] (sRGB straight colour, alpha 0/255)."""
    w = h = res
    xf = torch.tensor(np.asarray(ngp_xform, dtype=np.float32), device=device)
    v, u = torch.meshgrid((torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h,
                          (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w, indexing="ij")
    d_cam = torch.stack([(u - 0.5) * w / fx, (v - 0.5) * h / fy, torch.ones_like(u)], dim=-1)
    d = d_cam @ xf[:, :3].T
    d = d / d.norm(dim=-1, keepdim=True)
    o = xf[:, 3]
    best_t = torch.full((h, w), float("inf"), device=device)
    color = torch.zeros((h, w, 3), device=device)
    light = torch.tensor([0.35, 0.85, 0.40], device=device); light = light / light.norm()
    inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
    for (bmin, bmax, col) in boxes:
        bmin_t = torch.tensor(bmin, device=device); bmax_t = torch.tensor(bmax, device=device)
        t0 = (bmin_t - o) * inv; t1 = (bmax_t - o) * inv
        tn = torch.minimum(t0, t1); tf = torch.maximum(t0, t1)
        tnear, axis = tn.max(dim=-1); tfar = tf.min(dim=-1).values
        hit = (tnear < tfar) & (tnear > 0) & (tnear < best_t)
        n = torch.zeros_like(d)
        n.scatter_(-1, axis.unsqueeze(-1), -torch.sign(torch.gather(d, -1, axis.unsqueeze(-1))))
        shade = 0.35 + 0.65 * torch.clamp((n * light).sum(-1), min=0.0)
        c = torch.tensor(col, device=device) * shade.unsqueeze(-1)
        color = torch.where(hit.unsqueeze(-1), c, color)
        best_t = torch.where(hit, tnear, best_t)
    alpha = torch.isfinite(best_t)
    rgb8 = torch.clamp(torch.round(_linear_to_srgb(color) * 255.0), 0, 255).to(torch.uint8)
    rgb8 = torch.where(alpha.unsqueeze(-1), rgb8, torch.zeros_like(rgb8))
    a8 = (alpha.to(torch.uint8) * 255).unsqueeze(-1)
    return torch.cat([rgb8, a8], dim=-1).contiguous()


def make_lego_scene(n_images=100, res=800, device=None, seed=0, as_numpy=True):
    """Returns dict(images uint8 [n,res,res,4], xforms float32 [n,3,4] (ngp), fx, fy, cx, cy, aabb_scale)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    fx = fy = 0.5 * res / math.tan(0.5 * CAMERA_ANGLE_X)
    cams = hemisphere_cameras(n_images, seed=seed)
    xforms = np.stack([nerf_matrix_to_ngp(c) for c in cams]).astype(np.float32)
    boxes = lego_boxes()
    imgs = [render_image(xforms[i], res, fx, fy, boxes, device=device) for i in range(n_images)]
    images = torch.stack(imgs)
    if as_numpy:
        images = images.cpu().numpy()
    return dict(images=images, xforms=xforms, fx=float(fx), fy=float(fy), cx=0.5, cy=0.5, aabb_scale=1,
                nerf_c2w=[c.tolist() for c in cams], camera_angle_x=CAMERA_ANGLE_X, res=res)


def write_transforms_json(scene, out_dir, name="transforms_train.json"):
    """Writes a nerf_synthetic-style dataset (PNG + transforms json) the reference's loader would read."""
    from PIL import Image
    os.makedirs(os.path.join(out_dir, "train"), exist_ok=True)
    frames = []
    for i, c2w in enumerate(scene["nerf_c2w"]):
        rel = f"./train/r_{i}"
        Image.fromarray(np.asarray(scene["images"][i])).save(os.path.join(out_dir, rel + ".png"))
        frames.append({"file_path": rel, "transform_matrix": c2w})
    meta = {"camera_angle_x": scene["camera_angle_x"], "scale": NERF_SCALE, "offset": [NERF_OFFSET] * 3, "aabb_scale": scene["aabb_scale"], "frames": frames}
    path = os.path.join(out_dir, name)
    with open(path, "w") as f:
        json.dump(meta, f)
    return path
