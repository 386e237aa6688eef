"""Seeded synthetic inputs of SURVEY.md section 8(d): volumes, transfer-function control points, cameras, lights.

Input generation only (numpy); nothing here renders.
"""
import numpy as np

# ---- transfer functions as .tf1d control points: (rgb points [r,g,b,iso], alpha points [a,iso]) -----------------
# the one transfer function the reference ships (data/tf1dcp/bonsai_01.tf1d), as data
TF_BONSAI = (np.array([[0.00, 0.90, 0.00, 0], [0.00, 0.90, 0.00, 58], [0.00, 0.00, 0.00, 59],
                       [0.70, 0.35, 0.00, 60], [0.90, 0.55, 0.00, 255]], np.float64),
             np.array([[0.0, 0], [0.0, 35], [0.5, 40], [0.8, 255]], np.float64))
# TF-ramp: alpha 0@0 -> 0.8@255
TF_RAMP = (np.array([[0.1, 0.2, 0.9, 0], [0.9, 0.8, 0.1, 255]], np.float64),
           np.array([[0.0, 0], [0.8, 255]], np.float64))
# TF-sparse: alpha = 0 below iso 128
TF_SPARSE = (np.array([[0.2, 0.4, 1.0, 0], [1.0, 0.9, 0.3, 255]], np.float64),
             np.array([[0.0, 0], [0.0, 128], [0.6, 160], [0.9, 255]], np.float64))
# TF-zero: fully transparent (upper bound of the sample count, no termination)
TF_ZERO = (np.array([[1.0, 1.0, 1.0, 0], [1.0, 1.0, 1.0, 255]], np.float64),
           np.array([[0.0, 0], [0.0, 255]], np.float64))
# TF-thin: low opacity everywhere so that rays traverse the whole volume while every sample is shaded
TF_THIN = (np.array([[0.9, 0.3, 0.1, 0], [0.2, 0.8, 0.9, 255]], np.float64),
           np.array([[0.002, 0], [0.03, 255]], np.float64))
TFS = {"bonsai": TF_BONSAI, "ramp": TF_RAMP, "sparse": TF_SPARSE, "zero": TF_ZERO, "thin": TF_THIN}


def write_tf1d(path, tf, max_density=None, ext_flag=None):
    rgb, a = tf
    with open(path, "w") as f:
        f.write("linear\n")
        if max_density is None:
            f.write("0\n")
        elif ext_flag is None:
            f.write(f"1\n{int(max_density)}\n")
        else:
            f.write(f"2\n{int(max_density)} {int(ext_flag)}\n")
        f.write(f"{len(rgb)}\n")
        for r in rgb:
            f.write(f"{r[0]:.6f} {r[1]:.6f} {r[2]:.6f} {int(r[3])}\n")
        f.write(f"{len(a)}\n")
        for r in a:
            f.write(f"{r[0]:.6f} {int(r[1])}\n")


# ---- volumes: arrays indexed [z, y, x] -----------------------------------------------------------------------------
def volume_gauss(n, dtype=np.uint8):
    """The reference's own generator (libs/volvis_utils/utils.cpp:373-396): v = int(max * exp(-r^2 / 2 s^2)),
    centre n/2, s = n/4, fp32 arithmetic."""
    maxv = np.float32(255.0 if dtype == np.uint8 else 65535.0)
    s = np.float32(n / 4.0)
    c = np.arange(n, dtype=np.float32) - np.float32(n // 2)
    r2 = (c[:, None, None] ** 2 + c[None, :, None] ** 2 + c[None, None, :] ** 2).astype(np.float32)
    val = np.exp(-r2 / (np.float32(2.0) * s * s)).astype(np.float32)
    return (val * maxv).astype(np.int64).astype(dtype)


def volume_noise(n, dtype=np.uint8, seed=1234, octaves=3, base=8):
    """Band-limited value noise, 3 octaves, quantised (defeats empty-space skipping)."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    acc = np.zeros((n, n, n), np.float32)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        g = base * (2 ** o)
        coarse = rng.random((g + 1, g + 1, g + 1), dtype=np.float32)
        zoom = n / (g + 1)
        up = ndimage.zoom(coarse, zoom, order=1, mode="nearest", grid_mode=False)
        acc += amp * up[:n, :n, :n]
        tot += amp
        amp *= 0.5
    acc /= tot
    maxv = 255.0 if dtype == np.uint8 else 65535.0
    return np.clip(acc * maxv, 0, maxv).astype(dtype)


def volume_gauss_noise(n, dtype=np.uint16, seed=1234):
    """V-gauss modulated by V-noise (config 3)."""
    g = volume_gauss(n, np.uint16).astype(np.float32) / 65535.0
    nz = volume_noise(n, np.uint16, seed).astype(np.float32) / 65535.0
    maxv = 255.0 if dtype == np.uint8 else 65535.0
    return np.clip((0.65 * g + 0.35 * g * nz * 2.0) * maxv, 0, maxv).astype(dtype)


def boxes_records(n, count=64, seed=5678):
    """64 random axis-aligned boxes as .syn records (reader.cpp:307-321): (x0,y0,z0,x1,y1,z1,v), half-open."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(count):
        lo = rng.integers(0, n - n // 8, 3)
        sz = rng.integers(max(2, n // 32), max(3, n // 5), 3)
        hi = np.minimum(lo + sz, n)
        recs.append((int(lo[0]), int(lo[1]), int(lo[2]), int(hi[0]), int(hi[1]), int(hi[2]), int(rng.integers(30, 256))))
    return recs


def volume_boxes(n, count=64, seed=5678):
    v = np.zeros((n, n, n), np.uint8)
    for x0, y0, z0, x1, y1, z1, val in boxes_records(n, count, seed):
        v[z0:z1, y0:y1, x0:x1] = val
    return v


def write_syn_boxes(path, n, count=64, seed=5678):
    with open(path, "w") as f:
        f.write(f"{n} {n} {n}\n")
        for r in boxes_records(n, count, seed):
            f.write("1 " + " ".join(str(x) for x in r) + "\n")


# ---- cameras: the 10 states of the reference's data/#list_camera_states (for a 256^3 volume), scaled by n/256 -----
CAMERA_STATES_256 = [
    ("Initial State", (256.0, 256.0, 512.0), (0, 0, 0), (0, 1, 0)),
    ("Occlusion Comparison Synthetic", (-114.162, 75.4815, 102.521), (0, 0, 0), (0, 1, 0)),
    ("Occlusion Comparison VisMaleHead", (154.214, -298.045, -54.6573), (0, 0, 0), (0, 0, -1)),
    ("Occlusion Comparison Xmas Tree", (428.604, -467.801, 131.58), (0, 0, 0), (0, 0, 1)),
    ("Shadow Comparison Synthetic Bars", (131.959, 235.121, 335.252), (0, 0, 0), (0, 1, 0)),
    ("Shadow Comparison Engine", (-241.487, -128.506, -144.478), (0, 0, 0), (0, 0, -1)),
]
# first light of the first list of data/#list_light_sources
LIGHT0_256 = dict(position=(-206.873, -51.0699, 557.011), spot_angle=20.0)


def camera_state(idx, n):
    name, eye, center, up = CAMERA_STATES_256[idx]
    s = n / 256.0
    return tuple(e * s for e in eye), center, up


def light_position(n):
    s = n / 256.0
    return tuple(p * s for p in LIGHT0_256["position"])


def camera_forward(eye, center):
    """Camera::GetCameraVectors forward = -normalize(center - eye) (libs/vis_utils/camera.cpp:336-341); the reference
    copies it into the light's axes at start-up (renderingmanager.cpp:168,421-426)."""
    e = np.asarray(eye, np.float32); c = np.asarray(center, np.float32)
    d = c - e
    d = d / np.sqrt(np.sum(d * d, dtype=np.float32))
    return tuple((-d).tolist())


def volume_noise_torch(n, slices_zyx=None, dtype="u16", seed=1234, octaves=3, base=8, device="cuda"):
    """Band-limited value noise evaluated ON THE DEVICE for an arbitrary sub-block of an n^3 volume (sort-last bricks of
    volumes that do not fit the host, e.g. 2048^3): coarse seeded grids (identical on every rank) are interpolated
    separably at the block's voxel centres.  Returns a torch uint8 / int16-as-uint16-bits tensor [z, y, x] on `device`
    (uint16 values are stored in an int16 tensor's bits; pass .data_ptr() to vrb_volume_upload_device)."""
    import torch
    if slices_zyx is None:
        slices_zyx = (slice(0, n), slice(0, n), slice(0, n))
    rng = np.random.default_rng(seed)
    acc = None
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        g = base * (2 ** o)
        coarse = torch.from_numpy(rng.random((g + 1, g + 1, g + 1), dtype=np.float32)).to(device)
        cur = coarse
        for axis, sl in enumerate(slices_zyx):
            idx = torch.arange(sl.start, sl.stop, device=device, dtype=torch.float32)
            u = (idx + 0.5) / n * g
            i0 = torch.clamp(u.floor().long(), 0, g - 1)
            f = (u - i0.float()).clamp(0.0, 1.0)
            a = cur.index_select(axis, i0)
            b = cur.index_select(axis, i0 + 1)
            shape = [1, 1, 1]; shape[axis] = -1
            cur = a + (b - a) * f.view(shape)
        acc = cur * amp if acc is None else acc + cur * amp
        tot += amp
        amp *= 0.5
    acc = acc / tot
    if dtype == "u8":
        return (acc * 255.0).clamp(0, 255).to(torch.uint8).contiguous()
    v = (acc * 65535.0).clamp(0, 65535).to(torch.int32)
    return v.to(torch.uint16).contiguous() if hasattr(torch, "uint16") else (v - 65536 * (v >= 32768)).to(torch.int16).contiguous()


def make_volume(volume, dtype, n):
    """The seeded synthetic volumes of SURVEY.md section 8d by name: 'gauss', 'noise', 'gauss_noise', 'boxes'; dtype 'u8' / 'u16'."""
    dt = np.uint8 if dtype == "u8" else np.uint16
    if volume == "gauss":
        return volume_gauss(n, dt)
    if volume == "noise":
        return volume_noise(n, dt)
    if volume == "gauss_noise":
        return volume_gauss_noise(n, dt)
    if volume == "boxes":
        return volume_boxes(n)
    raise ValueError(volume)
