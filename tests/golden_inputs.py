"""Seeded inputs of the golden vectors under tests/golden/: shared by oracle/gen_golden.py (which feeds them to the
reference's own kernels on a B200 and stores the outputs) and by the tests (which feed them to the oracle / the CUDA path)."""
import numpy as np

N_GRID = 16384      # samples of the hash-grid forward / SH golden
N_GRID_BWD = 2048   # samples of the hash-grid backward golden (first N_GRID_BWD of the same positions)


def grid_inputs(n_grid_params, seed=1234):
    rs = np.random.RandomState(seed)
    table = (rs.randn(n_grid_params) * 0.5).astype(np.float16)
    positions = rs.rand(N_GRID, 3).astype(np.float32)
    positions[0] = 0.0; positions[1] = 1.0; positions[2] = [0.5, 0.25, 0.75]; positions[3] = [1.0, 0.0, 1.0]
    dy = (rs.randn(32, N_GRID) * 0.01).astype(np.float16)   # [feature][sample], the reference's layout
    dirs = rs.rand(N_GRID, 3).astype(np.float32)
    return table, positions, dy, dirs


N_MLP = 8192        # samples of the FullyFusedMLP golden


def mlp_inputs(n_hidden, seed=77):
    """Weights in FullyFusedMLP's parameter order ([64][32], (n_hidden-1) x [64][64], [16][64]), Xavier-like scale; inputs like hash-grid features /
    SH coefficients (|x| <= 1); output gradients at the magnitude the loss kernel produces (loss scale 128 / rays)."""
    rs = np.random.RandomState(seed + n_hidden)
    n_params = 64 * 32 + (n_hidden - 1) * 64 * 64 + 16 * 64
    w = (rs.uniform(-1, 1, n_params) * 0.25).astype(np.float16)
    x = (rs.uniform(-1, 1, (N_MLP, 32)) * 0.5).astype(np.float16)
    dy = np.zeros((N_MLP, 16), np.float16)
    dy[:, :4] = (rs.randn(N_MLP, 4) * 0.02).astype(np.float16)
    return w, x, dy


IMAGE_RES = 512     # BASELINE config 1: neural image 512 x 512, configs/image/base.json


def image_inputs(n_grid_params, seed=1337):
    """Fixed random parameters of the neural-image model (network 7168 halves, then the 2-D grid), and all pixel centres ((x+0.5)/W, (y+0.5)/H)."""
    rs = np.random.RandomState(seed)
    net = (rs.uniform(-1, 1, 7168) * 0.3).astype(np.float16)
    table = (rs.randn(n_grid_params) * 0.5).astype(np.float16)
    ys, xs = np.meshgrid(np.arange(IMAGE_RES), np.arange(IMAGE_RES), indexing="ij")
    uv = np.stack([(xs.reshape(-1) + 0.5) / IMAGE_RES, (ys.reshape(-1) + 0.5) / IMAGE_RES], axis=1).astype(np.float32)
    return net, table, uv


def image_grid_config():
    """configs/image/base.json with Testbed::reset_network's per_level_scale for a 512 x 512 image (desired_resolution = 256)."""
    pls = float(np.exp(np.log(np.float32(IMAGE_RES / 2.0) * np.float32(1) / np.float32(16)) / np.float32(15), dtype=np.float32))
    return dict(n_levels=16, log2_hashmap_size=24, base_resolution=16, per_level_scale=pls)


# ---- occupancy grid (K16): two cascades, inputs regenerated from seeds; only the reference kernels' outputs are stored ----
DG_CASCADES = 2
DG_CELLS = 128 ** 3 * DG_CASCADES
DG_SAMPLES = 16384
DG_STEP = 3
DG_AABB = np.array([-0.5, -0.5, -0.5, 1.5, 1.5, 1.5], np.float32)  # aabb_scale 2


def density_grid_inputs(seed=2024):
    """grid_in: 4 % of the cells above the occupancy threshold, 10 % marked untrained (-1), the rest small (the shape of a trained grid);
    density: fp16 network outputs (logits) for the DG_SAMPLES sampled cells."""
    rs = np.random.RandomState(seed)
    u = rs.rand(DG_CELLS)
    grid = (rs.rand(DG_CELLS) * 0.004).astype(np.float32)
    grid[u < 0.04] = (0.01 + rs.rand(int((u < 0.04).sum())) * 0.5).astype(np.float32)
    grid[u > 0.90] = -1.0
    density = (rs.randn(DG_SAMPLES) * 3.0).astype(np.float16)
    return np.ascontiguousarray(grid), density


def density_grid_cameras():
    """The 8-camera 64 x 64 synthetic scene of the other goldens (mark_untrained needs poses and intrinsics only)."""
    import synthetic
    scene = synthetic.make_lego_scene(8, 64, device="cpu", seed=0)
    return scene


# ---- K1 with the non-pinhole lens models (testbed_nerf.cu:1166-1190): (ELensMode, 7 parameters, principal point) applied to every camera of the 8 x 64^2 scene ----
LENS_CASES = {
    "opencv": (1, [0.0578421, -0.0805099, -0.000980296, 0.00015575, 0.0, 0.0, 0.0], (0.5135, 0.5027)),   # data/nerf/fox/transforms.json k1, k2, p1, p2, cx/w, cy/h
    "ftheta": (2, [0.0, 0.02, 1e-5, 0.0, 0.0, 64.0, 64.0], (0.5, 0.5)),                                     # alpha = 0.02 rad/px + 1e-5 rad/px^2 over a 64 x 64 sensor
    "latlong": (3, [0.0] * 7, (0.5, 0.5)),
}
LENS_N_RAYS, LENS_MAX_SAMPLES = 2048, 1 << 17


# ---- camera-extrinsics optimisation (K13 / K14) ----
N_CAM_SAMPLES = 4096
CAM_N_IMAGES, CAM_N_RAYS, CAM_N_KEPT = 8, 1024, 1000
CAM_AABB = np.array([-0.5, -0.5, -0.5, 1.5, 1.5, 1.5], np.float32)
CAM_ADAM_STEPS = 400


def camera_inputs(seed=4242):
    """Inputs of tests/golden/ref_camera.npz. Samples: positions / directions in [0,1], loss gradients at the hash features [n][32] and SH inputs [n][16].
    Rays: CAM_N_KEPT kept rays of CAM_N_RAYS with 0..12 compacted samples each; coords / coords_gradient [total][7]."""
    rs = np.random.RandomState(seed)
    n = N_CAM_SAMPLES
    d = dict(positions=rs.rand(n, 3).astype(np.float32), dirs=rs.rand(n, 3).astype(np.float32),
             dL_dencoded=(rs.randn(n, 32) * 0.01).astype(np.float16), dL_dsh=(rs.randn(n, 16) * 0.01).astype(np.float16))
    d["positions"][0] = 0.0; d["positions"][1] = 1.0; d["positions"][2] = [0.5, 0.25, 0.75]
    counts = rs.randint(0, 13, CAM_N_RAYS).astype(np.uint32)
    bases = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
    total = int(counts.sum())
    d["numsteps"] = np.stack([counts, bases], 1).astype(np.uint32)
    d["ray_indices"] = rs.randint(0, CAM_N_RAYS, CAM_N_RAYS).astype(np.uint32)
    rays = np.concatenate([rs.rand(CAM_N_RAYS, 3) * 2 - 0.5, rs.randn(CAM_N_RAYS, 3) * 3], 1).astype(np.float32)
    d["rays"] = rays
    coords = rs.rand(total, 7).astype(np.float32)
    d["coords"] = coords
    d["coords_gradient"] = (rs.randn(total, 7) * 0.05).astype(np.float32)
    # per-camera Adam: gradients spanning the magnitudes seen in training, the reference's learning-rate schedule (x200 for the rotation case so that the
    # composed rotations leave the small-angle regime)
    d["adam_gradients"] = (rs.randn(CAM_ADAM_STEPS, 3) * rs.choice([1e-3, 1.0, 30.0], size=(CAM_ADAM_STEPS, 1))).astype(np.float32)
    d["adam_lr"] = (1e-3 * 0.33 ** (np.arange(CAM_ADAM_STEPS) // 128)).astype(np.float32)
    d["offsets_xforms"] = rs.randn(6, 12).astype(np.float32)
    d["offsets_pos"] = (rs.randn(6, 3) * 0.1).astype(np.float32)
    d["offsets_rot"] = (rs.randn(6, 3) * 0.3).astype(np.float32)
    d["offsets_rot"][0] = 0.0
    return d


# ---- neural-image and SDF modes ----
MODE_IMAGE_RES = (512, 384)     # width, height: not square, so that the aspect handling of render_image is exercised
MODE_IMAGE_BATCH = 1 << 16
MODE_SDF_POOL = 1 << 16
MODE_SDF_BATCH = 1 << 14


def procedural_image(seed=5):
    """RGBA8 [h][w][4]: oriented sinusoids plus soft discs, opaque except for one transparent corner patch (exercises the alpha premultiplication)."""
    w, h = MODE_IMAGE_RES
    rs = np.random.RandomState(seed)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32) / h, np.arange(w, dtype=np.float32) / w, indexing="ij")
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(4):
            fx, fy, ph = rs.uniform(2, 24), rs.uniform(2, 24), rs.uniform(0, 6.28)
            img[..., c] += 0.12 * np.sin(6.2831853 * (fx * xs + fy * ys) + ph)
    for _ in range(12):
        cx, cy, r = rs.rand(), rs.rand(), rs.uniform(0.03, 0.15)
        col = rs.rand(3)
        m = np.clip(1.0 - np.hypot(xs - cx, ys - cy) / r, 0, 1)[..., None]
        img = img * (1 - m) + col * m
    img = np.clip(img + 0.5, 0, 1)
    alpha = np.ones((h, w, 1), np.float32)
    alpha[: h // 8, : w // 8] = 0.5
    return np.round(np.concatenate([img, alpha], -1) * 255.0).astype(np.uint8)


def sdf_pool(seed=9):
    """(positions [n][3] in the unit cube, distances [n]): the exact distance to the union of a sphere and a box (exact outside, a bound inside), uniform positions."""
    rs = np.random.RandomState(seed)
    n = MODE_SDF_POOL
    pos = rs.rand(n, 3).astype(np.float32)
    d_sphere = np.linalg.norm(pos - np.array([0.4, 0.5, 0.5], np.float32), axis=1) - 0.25
    q = np.abs(pos - np.array([0.65, 0.5, 0.5], np.float32)) - np.array([0.15, 0.2, 0.1], np.float32)
    d_box = np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(1), 0)
    return pos, np.minimum(d_sphere, d_box).astype(np.float32)


# ---- per-image exposures (K6 exposure term + the host-side exposure optimizer; src/testbed_nerf.cu:1403,:1558-1571,:3105-3131) ----
EXPOSURE_UPDATES = 200          # camera-update rounds of the host-side golden
EXPOSURE_N_IMAGES = 8
EXPOSURE_L2_REG = 1e-3
EXPOSURE_PER_CAMERA_LOSS_SCALE = 8.0 / 128.0 / 16.0  # n_images / LOSS_SCALE / n_steps_between_cam_updates


def exposure_inputs(n_images=EXPOSURE_N_IMAGES, seed=11):
    """exposures [n_images][3] in stops for the K6 golden; gradients [EXPOSURE_UPDATES][n_images][3] and learning rates for the optimizer golden."""
    rs = np.random.RandomState(seed)
    exposures = rs.uniform(-1.0, 1.0, (n_images, 3)).astype(np.float32)
    gradients = (rs.randn(EXPOSURE_UPDATES, n_images, 3) * np.exp(rs.uniform(-3, 3, (EXPOSURE_UPDATES, n_images, 1)))).astype(np.float32)
    learning_rates = np.where(np.arange(EXPOSURE_UPDATES) < 120, 1e-2, 3.3e-3).astype(np.float32)
    return dict(exposures=exposures, gradients=gradients, learning_rates=learning_rates)


# ---- end-to-end exposure optimisation (tests/golden/ref_exposure_train.npz: the reference's learned exposures on this dataset) ----
EXPOSURE_SCENE = dict(n_images=24, res=96, batch=1 << 16, steps=2000)


def exposure_scene_offsets(n_images=EXPOSURE_SCENE["n_images"], seed=21):
    """The per-image, per-channel exposure error (stops) the dataset is given: one offset per image plus a small per-channel part."""
    rs = np.random.RandomState(seed)
    return (rs.uniform(-0.6, 0.6, (n_images, 1)) + 0.1 * rs.uniform(-1, 1, (n_images, 3))).astype(np.float32)


def apply_image_exposures(images_u8, e):
    """images_u8 [n][h][w][4] sRGB straight-alpha; multiplies the LINEAR colours of image i by 2^e[i] and re-encodes (clipped) to 8-bit sRGB."""
    x = images_u8[..., :3].astype(np.float64) / 255.0
    lin = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
    lin = np.clip(lin * (2.0 ** e.astype(np.float64))[:, None, None, :], 0.0, 1.0)
    srgb = np.where(lin < 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055)
    out = images_u8.copy()
    out[..., :3] = np.clip(np.rint(srgb * 255.0), 0, 255).astype(np.uint8)
    return out


# ---- K19: error-map importance sampling (tests/golden/ref_error_map.npz) ----
ERROR_CDF_RES = (12, 10)   # (res_x, res_y) of the error map the CDFs are built from
ERROR_MAP_RES = (9, 11)    # (res_x, res_y) of the map K6 deposits into (the reference allows the two to differ: the map is re-sized at the start of a window)


def error_map_inputs(n_images, seed=31):
    """An accumulated error map [n_images][res_y][res_x] like a training window leaves it: per-ray losses ~1e-4 .. 1e-2 concentrated on a few texels, one
    row and one whole image without any error (the 1e-10 floor and the uniform blend keep those samplable), one image with ten times the error."""
    rx, ry = ERROR_CDF_RES
    rs = np.random.RandomState(seed)
    em = (rs.rand(n_images, ry, rx) ** 4 * 1e-2).astype(np.float32)
    em[:, 3, :] = 0.0
    em[1 % n_images] = 0.0
    em[2 % n_images] *= 10.0
    return em


# ---- end-to-end error-map sampling (tests/golden/ref_error_map_train.npz: the reference's window state and image probabilities on this dataset) ----
ERROR_SCENE = dict(n_images=8, res=64, batch=1 << 14, damaged=5, windows=(128, 192, 288))


def error_scene_images(images_u8):
    """The small scene with one image replaced by opaque mid-grey: no radiance field explains it together with the others, so its error stays high and the
    image CDF must favour it."""
    out = np.array(images_u8, copy=True)
    out[ERROR_SCENE["damaged"], ..., :3] = 128
    out[ERROR_SCENE["damaged"], ..., 3] = 255
    return out
