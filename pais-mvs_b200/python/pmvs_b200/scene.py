"""Scene construction for the patch-refinement path: cameras, grey/edge pyramids and the synthetic
multi-view workload the benchmarks and parity tests use.

Mirrors the reference's Camera (TMVS/mvs/camera.cpp:45-136): K,R,t from focal/principal/quaternion/centre,
KR = K*R, KT = K*t, optical normal = R^T*(0,0,1), a grey u8 pyramid (level i = INTER_AREA resize of level 0 by
lodRatio^i) and a min-max normalised gradient-magnitude f64 pyramid. Everything here is host-side numpy; the
arrays are handed to the C-ABI (pmvs_create) or to the oracle unchanged, so both see identical pyramids.
"""
import ctypes as C
import math

import numpy as np

from . import abi


# ---------------------------------------------------------------------------------------------------------
# camera math
# ---------------------------------------------------------------------------------------------------------
def quat_to_R(q):
    """Camera::quaternionToRotationMat, TMVS/mvs/camera.cpp:6-35 (q = w,x,y,z; normalised first)."""
    q = np.asarray(q, dtype=np.float64)
    qq = math.sqrt(float(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]))
    if qq > 0:
        qw, qx, qy, qz = (float(v) / qq for v in q)
    else:
        qw, qx, qy, qz = 1.0, 0.0, 0.0, 0.0
    R = np.empty((3, 3), dtype=np.float64)
    R[0, 0] = qw * qw + qx * qx - qz * qz - qy * qy
    R[0, 1] = 2 * qx * qy - 2 * qz * qw
    R[0, 2] = 2 * qy * qw + 2 * qz * qx
    R[1, 0] = 2 * qx * qy + 2 * qw * qz
    R[1, 1] = qy * qy + qw * qw - qz * qz - qx * qx
    R[1, 2] = 2 * qz * qy - 2 * qx * qw
    R[2, 0] = 2 * qx * qz - 2 * qy * qw
    R[2, 1] = 2 * qy * qz + 2 * qw * qx
    R[2, 2] = qz * qz + qw * qw - qy * qy - qx * qx
    return R


def R_to_quat(R):
    """Inverse of quat_to_R for a proper rotation (w >= 0 branch selection by largest diagonal term)."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        return np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    if R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    if R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        return np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
    return np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])


def camera_max_lod(cols, rows, cfg):
    """camera.cpp:63-64."""
    m = int(math.log(float(max(cols, rows))) / math.log(1.0 / cfg.lodRatio))
    return min(m, cfg.maxLOD)


# ---------------------------------------------------------------------------------------------------------
# pyramids (camera.cpp:63-92)
# ---------------------------------------------------------------------------------------------------------
def resize_area(img, f):
    """cv::resize(img, Size(), f, f, INTER_AREA) (camera.cpp:85) — by OpenCV itself: the synthetic scenes are built with the
    library the reference uses (cv2 here is 4.x; its INTER_AREA code path is the 2.4 one). The restatement the device
    kernels are checked against lives with the test infrastructure, not in this package."""
    import cv2
    return cv2.resize(np.ascontiguousarray(img), None, fx=f, fy=f, interpolation=cv2.INTER_AREA)


def edge_image(grey):
    """Sobel(ksize=1) gradient magnitude, min-max normalised (camera.cpp:71-78, 87-91), by OpenCV."""
    import cv2
    g = np.ascontiguousarray(grey)
    gx = cv2.Sobel(g, cv2.CV_64F, 1, 0, ksize=1)
    gy = cv2.Sobel(g, cv2.CV_64F, 0, 1, ksize=1)
    e = np.sqrt(gx * gx + gy * gy)
    mn, mx = e.min(), e.max()
    return np.ascontiguousarray((e - mn) / (mx - mn))


def build_pyramid(grey0, cfg, with_edge):
    """Level list [(grey u8, edge f64|None)], levels 0..maxLOD of this camera."""
    rows, cols = grey0.shape
    max_lod = camera_max_lod(cols, rows, cfg)
    levels = []
    for i in range(max_lod + 1):
        g = np.ascontiguousarray(grey0) if i == 0 else resize_area(grey0, cfg.lodRatio ** i)
        levels.append((g, edge_image(g) if with_edge else None))
    return levels


# ---------------------------------------------------------------------------------------------------------
# Camera record
# ---------------------------------------------------------------------------------------------------------
class Camera:
    """Host-side camera: numpy matrices + pyramid arrays; `record()` fills a PmvsCamera (borrowing the arrays)."""

    def __init__(self, focal, principal, quaternion, center, levels, name=""):
        self.name = name
        self.focal = np.array(focal, dtype=np.float64).reshape(2)
        self.principal = np.array(principal, dtype=np.float64).reshape(2)
        self.quaternion = np.array(quaternion, dtype=np.float64).reshape(4)
        self.center = np.array(center, dtype=np.float64).reshape(3)
        self.R = quat_to_R(self.quaternion)
        self.t = -(self.R @ self.center)                                   # camera.cpp:120
        K = np.array([[self.focal[0], 0, self.principal[0]], [0, self.focal[1], self.principal[1]], [0, 0, 1.0]])
        self.K = K
        self.KR = K @ self.R                                               # camera.cpp:123
        self.KT = K @ self.t                                               # camera.cpp:124
        self.optical_normal = self.R.T @ np.array([0.0, 0.0, 1.0])         # camera.cpp:130-133
        self.levels = levels
        self.max_lod = len(levels) - 1

    def project(self, X, lod=0, lod_ratio=0.8):
        """Camera::project without distortion, camera.cpp:138-160."""
        X2 = self.R @ np.asarray(X, dtype=np.float64) + self.t
        u = self.focal * (X2[:2] / X2[2]) + self.principal
        return u * (lod_ratio ** lod)

    def record(self):
        c = abi.PmvsCamera()
        for i in range(2):
            c.focal[i] = self.focal[i]
            c.principal[i] = self.principal[i]
        for i in range(3):
            c.center[i] = self.center[i]
            c.t[i] = self.t[i]
            c.KT[i] = self.KT[i]
            c.opticalNormal[i] = self.optical_normal[i]
        for i in range(9):
            c.R[i] = self.R.flat[i]
            c.KR[i] = self.KR.flat[i]
        c.maxLOD = self.max_lod
        for l, (g, e) in enumerate(self.levels):
            assert g.dtype == np.uint8 and g.flags["C_CONTIGUOUS"]
            c.level[l].cols = g.shape[1]
            c.level[l].rows = g.shape[0]
            c.level[l].pitch = g.strides[0]
            c.level[l].grey = g.ctypes.data
            if e is not None:
                assert e.dtype == np.float64 and e.flags["C_CONTIGUOUS"] and e.shape == g.shape
                c.level[l].edge = e.ctypes.data
            else:
                c.level[l].edge = None
        return c


def camera_array(cams):
    arr = (abi.PmvsCamera * len(cams))()
    for i, c in enumerate(cams):
        arr[i] = c.record()
    return arr


# ---------------------------------------------------------------------------------------------------------
# synthetic multi-view scene
# ---------------------------------------------------------------------------------------------------------
def _noise_texture(n, seed):
    from scipy.ndimage import gaussian_filter
    rng = np.random.RandomState(seed)
    t = gaussian_filter(rng.rand(n, n), sigma=2.0, mode="wrap")
    t = (t - t.mean()) / t.std()
    return t


def _look_at(center, target, up=(0.0, 1.0, 0.0)):
    z = np.asarray(target, dtype=np.float64) - center
    z /= np.linalg.norm(z)
    x = np.cross(np.asarray(up, dtype=np.float64), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])          # rows = camera axes in world coordinates: X_cam = R (X - C)


class SynthScene:
    """Textured plane z = plane_z seen by `nviews` cameras on an arc of +-arc_deg around the plane normal. (The plane
    is kept off the world origin: the reference's homography bracket d*K*R - K*t*n^T, patch.cpp:314, is singular
    for a plane through the origin, d = 0.)

    Images are band-limited noise rendered through the exact plane-to-image mapping, clamped to [1,255]
    (0 is the reference's background mask, patch.cpp:986). `background` > 0 paints that many zero-valued discs
    into every image to exercise the mask path."""

    def __init__(self, cfg, nviews=5, width=640, height=480, seed=1234, arc_deg=25.0, distance=10.0, with_edge=None,
                 background=0, tex_size=2048, plane_z=3.0):
        self.cfg = cfg
        self.plane_z = plane_z
        self.width, self.height = width, height
        self.distance = distance
        f = 1.2 * width
        self.focal = f
        tex = _noise_texture(tex_size, seed)
        self.texel = 0.75 * distance / f                 # world units per texel
        rng = np.random.RandomState(seed + 1)
        if with_edge is None:
            with_edge = bool(cfg.adaptiveGradientEnable)
        self.cams = []
        for k in range(nviews):
            a = 0.0 if nviews == 1 else math.radians(arc_deg) * (2.0 * k / (nviews - 1) - 1.0)
            b = math.radians(6.0) * math.sin(1.7 * k + 0.3)            # a little out-of-arc variation
            C0 = distance * np.array([math.sin(a) * math.cos(b), math.sin(b), math.cos(a) * math.cos(b)])
            C0[2] += plane_z
            R0 = _look_at(C0, np.array([0.0, 0.0, plane_z]))
            q = R_to_quat(R0)
            img = self._render(tex, quat_to_R(q), C0, f, width, height)
            if background:
                yy, xx = np.mgrid[0:height, 0:width]
                for _ in range(background):
                    cx, cy, rad = rng.randint(0, width), rng.randint(0, height), rng.randint(3, 12)
                    img[(xx - cx) ** 2 + (yy - cy) ** 2 <= rad * rad] = 0
            levels = build_pyramid(img, cfg, with_edge)
            self.cams.append(Camera((f, f), (width >> 1, height >> 1), q, C0, levels, name="synth%04d" % k))
        self.records = camera_array(self.cams)

    def _render(self, tex, R, Cw, f, width, height):
        n = tex.shape[0]
        u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
        d = np.stack([(u - (width >> 1)) / f, (v - (height >> 1)) / f, np.ones_like(u)], axis=-1)
        dw = d @ R                                      # R^T d, row-vector form
        t = (self.plane_z - Cw[2]) / dw[..., 2]
        X = Cw[0] + t * dw[..., 0]
        Y = Cw[1] + t * dw[..., 1]
        tx = X / self.texel + n / 2.0
        ty = Y / self.texel + n / 2.0
        x0 = np.floor(tx)
        y0 = np.floor(ty)
        fx = tx - x0
        fy = ty - y0
        x0 = x0.astype(np.int64) % n
        y0 = y0.astype(np.int64) % n
        x1 = (x0 + 1) % n
        y1 = (y0 + 1) % n
        val = (tex[y0, x0] * (1 - fx) * (1 - fy) + tex[y0, x1] * fx * (1 - fy) + tex[y1, x0] * (1 - fx) * fy +
               tex[y1, x1] * fx * fy)
        img = np.clip(np.rint(128.0 + 110.0 * val), 1, 255).astype(np.uint8)
        return np.ascontiguousarray(img)

    # -- patch candidates ---------------------------------------------------------------------------------
    def patches(self, n, seed=5678, ptype=abi.TYPE_EXPAND, normal_jitter_deg=10.0, depth_jitter=0.04, extent=None,
                first_id=0):
        """n candidate patches on a jittered grid over the part of the plane every camera sees: centre
        perturbed along the central viewing ray, normal = truth (0,0,1) perturbed by <= normal_jitter_deg,
        camIdx = all views (what the expansion ctor hands refine() after expandVisibleCamera)."""
        rng = np.random.RandomState(seed)
        if extent is None:
            # half-size of the plane region that projects well inside every view (window + parallax margin)
            extent = 0.30 * self.distance * min(self.width, self.height) / self.focal
        side = int(math.ceil(math.sqrt(n)))
        arr = (abi.PmvsPatchIn * n)()
        if n == 0:
            return arr
        nviews = len(self.cams)
        r = rng.rand(n, 5)
        k = np.arange(n)
        X = (-1 + 2 * ((k % side) + r[:, 0]) / side) * extent
        Y = (-1 + 2 * ((k // side) + r[:, 1]) / side) * extent
        center = np.stack([X, Y, np.full(n, self.plane_z)], axis=1)
        ray = center - np.array([0.0, 0.0, self.plane_z + self.distance])
        ray /= np.linalg.norm(ray, axis=1, keepdims=True)
        center = center + ray * ((2 * r[:, 2:3] - 1) * depth_jitter)
        tilt = math.radians(normal_jitter_deg) * r[:, 3]
        az = 2 * math.pi * r[:, 4]
        nrm = np.stack([np.sin(tilt) * np.cos(az), np.sin(tilt) * np.sin(az), np.cos(tilt)], axis=1)
        v = np.frombuffer(arr, dtype=PATCH_IN_DTYPE)
        v["center"] = center
        v["normal"] = nrm
        v["normalS"][:, 0] = np.arccos(nrm[:, 2])                 # utility.h:17-22
        v["normalS"][:, 1] = np.arctan2(nrm[:, 1], nrm[:, 0])
        v["type"] = ptype
        v["id"] = first_id + k
        v["nCam"] = nviews
        v["camIdx"][:, :nviews] = np.arange(nviews, dtype=np.uint16)
        return arr


PATCH_IN_DTYPE = np.dtype([("center", "<f8", 3), ("normal", "<f8", 3), ("normalS", "<f8", 2), ("type", "<i4"), ("id", "<i4"),
                           ("nCam", "<i4"), ("_pad", "<i4"), ("camIdx", "<u2", abi.MAX_VIEWS)])
PATCH_OUT_DTYPE = np.dtype([("center", "<f8", 3), ("normal", "<f8", 3), ("normalS", "<f8", 2), ("ray", "<f8", 3), ("depth", "<f8"),
                            ("depthRange", "<f8", 2), ("fitness", "<f8"), ("priority", "<f8"), ("correlation", "<f8"),
                            ("LOD", "<i4"), ("refCamIdx", "<i4"), ("nCam", "<i4"), ("drop", "<i4"), ("psoRuns", "<i4"),
                            ("psoIterations", "<i4"), ("evaluations", "<u4"), ("status", "<u4"),
                            ("camIdx", "<u2", abi.MAX_VIEWS), ("nImgPoint", "<i4"), ("windowEvaluations", "<u4"),
                            ("imgPoint", "<f8", (abi.MAX_VIEWS, 2))])
assert PATCH_IN_DTYPE.itemsize == C.sizeof(abi.PmvsPatchIn) and PATCH_OUT_DTYPE.itemsize == C.sizeof(abi.PmvsPatchOut)


def hypotheses_from_patches(scene, patches, cfg, lod=0, seed=99, per_patch=4, spread=1.0):
    """Seam-1 inputs: for each patch a few (theta, phi, depth) hypotheses around its state, with the reference
    camera / ray refine() would pick (patch.cpp:415-461)."""
    rng = np.random.RandomState(seed)
    n = len(patches) * per_patch
    arr = (abi.PmvsHypothesis * n)()
    k = 0
    for p in patches:
        nrm = np.array(p.normal[:])
        best, ref = -1e300, -1
        for i in range(p.nCam):
            c = float(nrm @ (-scene.cams[p.camIdx[i]].optical_normal))
            if c > best:
                best, ref = c, p.camIdx[i]
        Cc = scene.cams[ref].center
        ray = np.array(p.center[:]) - Cc
        depth = float(np.linalg.norm(ray))
        ray = ray / depth
        for j in range(per_patch):
            h = arr[k]
            for i in range(3):
                h.ray[i] = ray[i]
            s = 0.0 if j == 0 else spread
            h.theta = p.normalS[0] + s * 0.3 * (rng.rand() - 0.5)
            h.phi = p.normalS[1] + s * 0.6 * (rng.rand() - 0.5)
            h.depth = depth + s * 0.1 * (rng.rand() - 0.5)
            h.refCamIdx = ref
            h.LOD = lod
            h.nCam = p.nCam
            for i in range(p.nCam):
                h.camIdx[i] = p.camIdx[i]
            k += 1
    return arr
