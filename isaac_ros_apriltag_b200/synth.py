"""Seeded synthetic frames for the BASELINE.json configs.

Tags are rendered from the code tables with exact geometry: each tag is a plane with a pose (R, t) in
front of a pinhole camera K; the cell image is point-sampled at 4x4 sub-pixel positions through the exact
homography and box-filtered, so the ground-truth corners are exact in AprilTag's pixel convention (pixel
(i, j) covers [i, i+1) x [j, j+1), centre at +0.5).
"""
import math

import cv2
import numpy as np

from .families import families, tag_cells

SS = 4  # supersampling factor per axis


def pose_homography(K, R, t, tagsize):
    """tag coords [-1,1]^2 (x right, y down on the tag, z into the tag) -> pixels."""
    s = tagsize / 2.0
    M = np.column_stack([R[:, 0] * s, R[:, 1] * s, t])
    return K @ M


def project(H, pts):
    p = np.asarray(pts, float)
    q = (H @ np.column_stack([p, np.ones(len(p))]).T).T
    return q[:, :2] / q[:, 2:3]


def render_tag(img, family, tag_id, H, black=20, white=235):
    """Composite one tag (including its 1-cell white quiet ring) into float32 `img` in place."""
    fam = families()[family]
    tw, wab = fam["total_width"], fam["width_at_border"]
    a = tw / wab  # tag-coordinate half extent of the full tag incl. white ring
    cells = tag_cells(family, tag_id)
    tex = np.where(cells > 0, np.float32(white), np.float32(black))
    outer = project(H, [(-a, -a), (a, -a), (a, a), (-a, a)])
    x0 = max(int(math.floor(outer[:, 0].min())) - 1, 0)
    y0 = max(int(math.floor(outer[:, 1].min())) - 1, 0)
    x1 = min(int(math.ceil(outer[:, 0].max())) + 1, img.shape[1])
    y1 = min(int(math.ceil(outer[:, 1].max())) + 1, img.shape[0])
    if x1 <= x0 or y1 <= y0:
        return
    pw, ph = (x1 - x0) * SS, (y1 - y0) * SS
    # dst (supersampled patch index) -> frame coords -> tag coords -> texel index
    A1 = np.array([[1.0 / SS, 0, x0 + 0.5 / SS], [0, 1.0 / SS, y0 + 0.5 / SS], [0, 0, 1]])
    T = np.array([[tw / (2 * a), 0, tw / 2.0 - 0.5], [0, tw / (2 * a), tw / 2.0 - 0.5], [0, 0, 1]])
    Minv = T @ np.linalg.inv(H) @ A1
    flags = cv2.INTER_NEAREST | cv2.WARP_INVERSE_MAP
    patch = cv2.warpPerspective(tex, Minv, (pw, ph), flags=flags, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    mask = cv2.warpPerspective(np.ones_like(tex), Minv, (pw, ph), flags=flags, borderMode=cv2.BORDER_CONSTANT,
                               borderValue=0)
    patch = cv2.resize(patch, (x1 - x0, y1 - y0), interpolation=cv2.INTER_AREA)
    mask = cv2.resize(mask, (x1 - x0, y1 - y0), interpolation=cv2.INTER_AREA)
    roi = img[y0:y1, x0:x1]
    roi[:] = roi * (1 - mask) + patch


def default_K(width, height):
    f = 1000.0 * width / 1920.0
    return np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])


def _rot(axis, ang):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * Kx + (1 - math.cos(ang)) * (Kx @ Kx)


def background(rng, width, height, noise_sigma):
    # low-frequency gradient 96..160
    gx = rng.uniform(-1, 1)
    gy = rng.uniform(-1, 1)
    xs = np.linspace(-1, 1, width, dtype=np.float32)[None, :]
    ys = np.linspace(-1, 1, height, dtype=np.float32)[:, None]
    g = gx * xs + gy * ys
    g = (g - g.min()) / max(g.max() - g.min(), 1e-6)
    return (96 + 64 * g).astype(np.float32)


def make_frame(rng, width, height, tags, K=None, tagsize=0.22, side_px=(48, 220), max_tilt_deg=45.0, noise_sigma=2.0,
               blur_sigma=0.7, margin=8, grid=None, black=20, white=235):
    """tags: list of (family, id).  Returns (gray uint8 HxW, ground truth list).

    grid: optional (cols, rows, pitch_px, side_px) -> fronto-parallel grid layout (config C4)."""
    if K is None:
        K = default_K(width, height)
    img = background(rng, width, height, noise_sigma)
    truth = []
    boxes = []
    if grid is not None:
        cols, rows, pitch, side = grid
        x_off = (width - cols * pitch) / 2.0 + pitch / 2.0
        y_off = (height - rows * pitch) / 2.0 + pitch / 2.0
    for ti, (family, tag_id) in enumerate(tags):
        fam = families()[family]
        a = fam["total_width"] / fam["width_at_border"]
        for _attempt in range(200):
            if grid is not None:
                cxp = x_off + (ti % cols) * pitch + rng.uniform(-0.5, 0.5)
                cyp = y_off + (ti // cols) * pitch + rng.uniform(-0.5, 0.5)
                z = K[0, 0] * tagsize / side
                R = _rot([0, 0, 1], rng.integers(0, 4) * math.pi / 2 + rng.uniform(-0.05, 0.05))
            else:
                side = rng.uniform(*side_px)
                z = K[0, 0] * tagsize / side
                cxp = rng.uniform(margin + side * 0.8, width - margin - side * 0.8)
                cyp = rng.uniform(margin + side * 0.8, height - margin - side * 0.8)
                inplane = rng.uniform(0, 2 * math.pi)
                tilt = math.radians(rng.uniform(0, max_tilt_deg))
                tax = rng.uniform(0, 2 * math.pi)
                R = _rot([math.cos(tax), math.sin(tax), 0], tilt) @ _rot([0, 0, 1], inplane)
            t = np.array([(cxp - K[0, 2]) / K[0, 0] * z, (cyp - K[1, 2]) / K[1, 1] * z, z])
            H = pose_homography(K, R, t, tagsize)
            outer = project(H, [(-a, -a), (a, -a), (a, a), (-a, a)])
            bx0, by0 = outer.min(0)
            bx1, by1 = outer.max(0)
            if bx0 < margin or by0 < margin or bx1 > width - margin or by1 > height - margin:
                continue
            if any(not (bx1 + 4 < o[0] or o[2] + 4 < bx0 or by1 + 4 < o[1] or o[3] + 4 < by0) for o in boxes):
                if grid is None:
                    continue
            boxes.append((bx0, by0, bx1, by1))
            render_tag(img, family, tag_id, H, black, white)
            truth.append({"family": family, "id": int(tag_id), "H": H / H[2, 2], "R": R, "t": t,
                          "center": project(H, [(0, 0)])[0],
                          # AprilRobotics order: (-1,1),(1,1),(1,-1),(-1,-1)
                          "p": project(H, [(-1, 1), (1, 1), (1, -1), (-1, -1)])})
            break
    if blur_sigma and blur_sigma > 0:
        img = cv2.GaussianBlur(img, (0, 0), blur_sigma)
    if noise_sigma and noise_sigma > 0:
        img = img + rng.normal(0, noise_sigma, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8), truth


def make_config_frames(config, n_frames, seed=None):
    """Seeded frames for BASELINE.json configs.  config in {"C1","C2","C3","C4","C5"}.
    Returns (frames (n,h,w) uint8, truths, K, tagsize, families)."""
    spec = {
        "C1": dict(w=1280, h=720, seed=1, tags=[("tag36h11", 1)], fams=("tag36h11",)),
        "C2": dict(w=1920, h=1080, seed=2, tags=[("tag36h11", 10)], fams=("tag36h11",)),
        "C3": dict(w=3840, h=2160, seed=3, tags=[("tag36h11", 4)], fams=("tag36h11",)),
        "C4": dict(w=1280, h=720, seed=4, tags=[("tag36h11", 91)], fams=("tag36h11",), grid=(13, 7, 96, 64)),
        "C5": dict(w=1920, h=1080, seed=5, tags=[("tag36h11", 6), ("tag25h9", 4)], fams=("tag36h11", "tag25h9")),
    }[config]
    rng = np.random.default_rng(spec["seed"] if seed is None else seed)
    w, h = spec["w"], spec["h"]
    K = default_K(w, h)
    frames = np.empty((n_frames, h, w), np.uint8)
    truths = []
    for i in range(n_frames):
        tags = []
        for fam, cnt in spec["tags"]:
            ids = rng.choice(len(families()[fam]["codes"]), size=cnt, replace=False)
            tags += [(fam, int(t)) for t in ids]
        kw = {}
        if "grid" in spec:
            kw["grid"] = spec["grid"]
        if config == "C3":
            kw["side_px"] = (96, 440)
        frames[i], tr = make_frame(rng, w, h, tags, K=K, **kw)
        truths.append(tr)
    return frames, truths, K, 0.22, spec["fams"]


def make_apriltag0():
    """Re-synthesis of the reference's only fixture (isaac_ros_apriltag/test/test_cases/apriltag0, an
    unmaterialised LFS pointer): 1920x1080 bgr8, tag36h11 id 0, 236 px, rotated 180 deg in plane, border
    spanning x in [808,1044], y in [429,665] (derivation: SURVEY.md section 4)."""
    img = np.full((1080, 1920), 255, np.float32)
    H = np.array([[-118.0, 0, 926.0], [0, -118.0, 547.0], [0, 0, 1.0]])
    render_tag(img, "tag36h11", 0, H, black=0, white=255)
    gray = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    bgr = np.repeat(gray[:, :, None], 3, axis=2)
    K = np.array([[434.943999, 0, 651.073921], [0, 431.741273, 441.878037], [0, 0, 1.0]])
    return bgr, K
