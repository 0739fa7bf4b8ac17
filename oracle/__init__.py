"""ctypes binding of the CPU oracle (oracle/pft_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from pcl_tracking_b200/ (the product path).
PARITY UNPINNED: PCL 1.8.0 (where the reference's arithmetic lives) is not vendored in
/root/reference and not installable here; see the header of pft_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpft_oracle.so")

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
PARTICLE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("one", "<f4"),
                     ("roll", "<f4"), ("pitch", "<f4"), ("yaw", "<f4"), ("weight", "<f4")])

# keys (mirror the enum in pft_oracle.cpp)
THREADS, PARTICLE_NUM, MAX_PARTICLE_NUM, ITERATION_NUM, NN_MODE, USE_HSV, USE_DISTANCE, SAMPLER, QUAT_SAMPLE, SEED = range(10)
DELTA, EPSILON, ALPHA, MOTION_RATIO, MAX_DIST, DIST_WEIGHT, HSV_WEIGHT, H_WEIGHT, S_WEIGHT, V_WEIGHT, OCTREE_RES, GRID_CELL = range(20, 32)
STEP_COV, INIT_COV, INIT_MEAN, BIN_SIZE = range(40, 44)
USE_CHANGE_DETECTOR, CD_INTERVAL, CD_MIN_POINTS = 10, 11, 12
CD_RESOLUTION = 32
NN_EXACT_BRUTE, NN_EXACT_GRID, NN_PCL_APPROX = 0, 1, 2
SAMPLER_ALIAS_PCL, SAMPLER_CDF, SAMPLER_CDF_VDC = 0, 1, 2


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (g++ only, a few seconds)."""
    src = os.path.join(_HERE, "pft_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
        sig = {
            "orc_passthrough": (i, [vp, i, i, f, f, vp]),
            "orc_approx_voxel_grid_pcl": (i, [vp, i, f, vp]),
            "orc_voxel_grid_pcl": (i, [vp, i, f, vp]),
            "orc_voxel_grid_exact": (i, [vp, i, f, i, f, f, vp]),
            "orc_remove_zero_points": (i, [vp, i, vp]),
            "orc_euclidean_clusters": (i, [vp, i, d, i, i, vp, vp, i]),
            "orc_centroid": (None, [vp, i, vp]),
            "orc_rgb2hsv": (None, [i, i, i, vp, vp, vp]),
            "orc_div_table": (i, [i]),
            "orc_normal_quantile": (d, [d]),
            "orc_kl_bound": (d, [i, d, d]),
            "orc_particle_to_matrix": (None, [vp, vp]),
            "orc_matrix_to_particle": (None, [vp, vp]),
            "orc_distance_coherence": (d, [vp, vp, d]),
            "orc_hsv_coherence": (d, [C.c_uint32, C.c_uint32, d, d, d, d]),
            "orc_particle_sample": (None, [vp, vp, vp, vp, i]),
            "orc_octree_approx_nearest": (i, [vp, i, d, vp, i, vp, vp]),
            "orc_tracker_create": (vp, [i]),
            "orc_tracker_destroy": (None, [vp]),
            "orc_set_i": (i, [vp, i, i]),
            "orc_set_d": (i, [vp, i, d]),
            "orc_set_vec6": (i, [vp, i, vp]),
            "orc_set_trans": (None, [vp, vp]),
            "orc_set_reference": (None, [vp, vp, i]),
            "orc_set_input": (None, [vp, vp, i]),
            "orc_set_particles": (None, [vp, vp, i]),
            "orc_get_particles": (i, [vp, vp, i]),
            "orc_get_result": (None, [vp, vp]),
            "orc_set_result": (None, [vp, vp]),
            "orc_get_motion": (None, [vp, vp]),
            "orc_set_motion": (None, [vp, vp]),
            "orc_set_changed": (None, [vp, i]),
            "orc_reset_tracking": (None, [vp]),
            "orc_alias_table": (None, [vp, vp, vp]),
            "orc_get_changed": (i, [vp]),
            "orc_change_detector_info": (None, [vp, vp]),
            "orc_change_detector_sequence": (i, [vp, vp, i, d, i, vp]),
            "orc_inject_draws": (None, [vp, vp, vp, vp, i, i]),
            "orc_init_particles": (None, [vp]),
            "orc_resample": (None, [vp, i]),
            "orc_weight": (None, [vp, i]),
            "orc_normalize": (None, [vp]),
            "orc_update": (None, [vp]),
            "orc_compute": (None, [vp]),
            "orc_get_aabb": (None, [vp, vp]),
            "orc_get_local_aabb": (None, [vp, vp]),
            "orc_set_crop_box": (None, [vp, vp]),
            "orc_get_cropped": (i, [vp, vp, vp, i]),
            "orc_get_nn": (i, [vp, i, vp, vp, i]),
            "orc_get_raw_weights": (i, [vp, vp, i]),
            "orc_get_ancestors": (i, [vp, vp, i]),
            "orc_get_fit_ratio": (d, [vp]),
            "orc_get_stage_seconds": (None, [vp, vp, i]),
            "orc_max_threads": (i, []),
            "orc_describe": (C.c_char_p, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def as_points(a):
    a = np.ascontiguousarray(a, dtype=POINT)
    return a


def make_points(xyz, rgba=None):
    xyz = np.asarray(xyz, dtype=np.float32)
    out = np.zeros(len(xyz), dtype=POINT)
    out["x"], out["y"], out["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if rgba is not None:
        out["rgba"] = np.asarray(rgba, dtype=np.uint32)
    return out


def make_particles(states, weights=None):
    s = np.asarray(states, dtype=np.float32).reshape(-1, 6)
    out = np.zeros(len(s), dtype=PARTICLE)
    for k, name in enumerate(("x", "y", "z", "roll", "pitch", "yaw")):
        out[name] = s[:, k]
    out["one"] = 1.0
    if weights is not None:
        out["weight"] = np.asarray(weights, dtype=np.float32)
    return out


# ------------------------------------------------------------------ free functions
def passthrough(pts, field, lo, hi):
    pts = as_points(pts)
    out = np.empty_like(pts)
    n = lib().orc_passthrough(_p(pts), len(pts), field, lo, hi, _p(out))
    return out[:n].copy()


def approx_voxel_grid_pcl(pts, leaf):
    pts = as_points(pts)
    out = np.empty(max(len(pts), 1), dtype=POINT)
    n = lib().orc_approx_voxel_grid_pcl(_p(pts), len(pts), leaf, _p(out))
    return out[:n].copy()


def voxel_grid_pcl(pts, leaf):
    pts = as_points(pts)
    out = np.empty(max(len(pts), 1), dtype=POINT)
    n = lib().orc_voxel_grid_pcl(_p(pts), len(pts), leaf, _p(out))
    return out[:n].copy()


def voxel_grid_exact(pts, leaf, field=2, lo=0.0, hi=10.0):
    pts = as_points(pts)
    out = np.empty(max(len(pts), 1), dtype=POINT)
    n = lib().orc_voxel_grid_exact(_p(pts), len(pts), leaf, field, lo, hi, _p(out))
    return out[:n].copy()


def from_pointcloud2(data, width, height, point_step, row_step=None, off_x=0, off_y=4, off_z=8, off_rgb=16):
    """pcl::fromPCLPointCloud2 for PointXYZRGBA (PCL-1.8.0 common/include/pcl/conversions.h: one memcpy per mapped field
    and point, rows of row_step bytes, records of point_step bytes; call site ref: src/auto_tracking.cpp:619-622), restated
    with numpy byte views.  Returns the packed {x,y,z,rgba} points in row-major order; NaNs are kept."""
    row_step = width * point_step if row_step is None else row_step
    raw = np.frombuffer(bytes(data), dtype=np.uint8)[:row_step * height].reshape(height, row_step)
    rec = raw[:, :width * point_step].reshape(height * width, point_step)
    out = np.zeros(height * width, dtype=POINT)
    for name, off in (("x", off_x), ("y", off_y), ("z", off_z)):
        out[name] = np.ascontiguousarray(rec[:, off:off + 4]).view("<f4")[:, 0]
    if off_rgb >= 0:
        out["rgba"] = np.ascontiguousarray(rec[:, off_rgb:off_rgb + 4]).view("<u4")[:, 0]
    return out


def euclidean_clusters(pts, tolerance=0.02, min_size=500, max_size=25000):
    """pcl::EuclideanClusterExtraction (ref: src/create_model.cpp:169-179).  Returns (labels per point: cluster rank
    or -1, cluster sizes in rank order: descending)."""
    pts = as_points(pts)
    labels = np.full(len(pts), -1, dtype=np.int32)
    sizes = np.zeros(max(len(pts), 1), dtype=np.int32)
    k = lib().orc_euclidean_clusters(_p(pts), len(pts), float(tolerance), int(min_size), int(max_size), _p(labels), _p(sizes), len(sizes))
    return labels, sizes[:k].copy()


def segment_plane(pts, samples3, distance_threshold=0.015, max_iterations=1000, probability=0.99, optimize=True):
    """pcl::SACSegmentation(SACMODEL_PLANE, SAC_RANSAC)::segment (ref: src/create_model_planar_segmentation.cpp:157-163)
    restated with numpy from PCL-1.8.0 sample_consensus/impl/ransac.hpp (computeModel), impl/sac_model_plane.hpp
    (computeModelCoefficients, countWithinDistance, selectWithinDistance, optimizeModelCoefficients) and
    segmentation/impl/sac_segmentation.hpp, with the RANSAC draws injected (`samples3`: rows of three point indices,
    consumed in order; a collinear sample is skipped without counting as an iteration -- upstream redraws).
    fp32 arithmetic where upstream uses Eigen float vectors (4-lane reductions add as (0 + 2) + (1 + 3)); the
    refinement accumulates in fp64 (upstream: sequential fp32).  Returns dict(coeff, ransac_coeff, best, iterations, inliers mask)."""
    f = np.float32
    pts = as_points(pts)
    x, y, z = pts["x"].astype(f), pts["y"].astype(f), pts["z"].astype(f)
    n = len(pts)

    def dist_mask(c):
        with np.errstate(invalid="ignore", over="ignore"):
            d = ((c[0] * x + c[2] * z).astype(f) + (c[1] * y + c[3]).astype(f)).astype(f)
            return np.abs(d).astype(np.float64) < distance_threshold

    def model(i0, i1, i2):
        if not (0 <= i0 < n and 0 <= i1 < n and 0 <= i2 < n) or len({i0, i1, i2}) < 3:
            return None
        with np.errstate(all="ignore"):
            p0 = np.array([x[i0], y[i0], z[i0]], dtype=f)
            a = (np.array([x[i1], y[i1], z[i1]], dtype=f) - p0).astype(f)
            b = (np.array([x[i2], y[i2], z[i2]], dtype=f) - p0).astype(f)
            r = (a / b).astype(f)
            if r[0] == r[1] and r[2] == r[1]:
                return None
            nrm = np.array([f(a[1] * b[2]) - f(a[2] * b[1]), f(a[2] * b[0]) - f(a[0] * b[2]), f(a[0] * b[1]) - f(a[1] * b[0])], dtype=f)
            ln = np.sqrt(f(f(nrm[0] * nrm[0]) + f(nrm[2] * nrm[2])) + f(f(nrm[1] * nrm[1]) + f(0)), dtype=f)
            nrm = (nrm / ln).astype(f)
            d = f(-1.0) * f(f(f(nrm[0] * p0[0]) + f(nrm[2] * p0[2])) + f(f(nrm[1] * p0[1]) + f(0)))
            c = np.array([nrm[0], nrm[1], nrm[2], d], dtype=f)
        return c if np.isfinite(c).all() else None

    iterations, best, best_count, skipped, k = 0, -1, -(2 ** 31 - 1), 0, 1.0
    log_p = np.log(1.0 - probability)
    best_c = None
    for s, (i0, i1, i2) in enumerate(np.asarray(samples3).reshape(-1, 3).tolist()):
        if not (iterations < k and skipped < max_iterations * 10):
            break
        c = model(i0, i1, i2)
        if c is None:
            skipped += 1
            continue
        cnt = int(dist_mask(c).sum())
        if cnt > best_count:
            best_count, best, best_c = cnt, s, c
            w = cnt / float(n)
            p_no = min(max(1.0 - w * w * w, np.finfo(np.float64).eps), 1.0 - np.finfo(np.float64).eps)
            k = log_p / np.log(p_no)
        iterations += 1
        if iterations > max_iterations:
            break
    if best < 0:
        return dict(coeff=None, ransac_coeff=None, best=-1, iterations=iterations, inliers=np.zeros(n, dtype=bool))
    coeff = best_c.copy()
    m = dist_mask(best_c)
    if optimize and m.sum() >= 3:
        q = np.stack([x[m], y[m], z[m]], 1)
        xx = np.stack([q[:, 0] * q[:, 0], q[:, 0] * q[:, 1], q[:, 0] * q[:, 2], q[:, 1] * q[:, 1], q[:, 1] * q[:, 2], q[:, 2] * q[:, 2]], 1).astype(f)
        a6 = xx.astype(np.float64).sum(0) / m.sum()
        mean = q.astype(np.float64).sum(0) / m.sum()
        cov = np.array([[a6[0], a6[1], a6[2]], [a6[1], a6[3], a6[4]], [a6[2], a6[4], a6[5]]]) - np.outer(mean, mean)
        w_, v_ = np.linalg.eigh(cov)
        nrm = v_[:, 0]
        if nrm @ best_c[:3].astype(np.float64) < 0:
            nrm = -nrm
        nf, mf = nrm.astype(f), mean.astype(f)
        d = f(-1.0) * f(f(f(nf[0] * mf[0]) + f(nf[1] * mf[1])) + f(nf[2] * mf[2]))
        coeff = np.array([nf[0], nf[1], nf[2], d], dtype=f)
        m = dist_mask(coeff)
    return dict(coeff=coeff, ransac_coeff=best_c, best=best, iterations=iterations, inliers=m)


def remove_zero_points(pts):
    pts = as_points(pts)
    out = np.empty_like(pts)
    n = lib().orc_remove_zero_points(_p(pts), len(pts), _p(out))
    return out[:n].copy()


def centroid(pts):
    pts = as_points(pts)
    c = np.zeros(3, dtype=np.float32)
    lib().orc_centroid(_p(pts), len(pts), _p(c))
    return c


def rgb2hsv(r, g, b):
    h, s, v = C.c_int(), C.c_int(), C.c_int()
    lib().orc_rgb2hsv(r, g, b, C.byref(h), C.byref(s), C.byref(v))
    return h.value, s.value, v.value


def div_table(i):
    return lib().orc_div_table(i)


def normal_quantile(u):
    return lib().orc_normal_quantile(u)


def kl_bound(k, delta, eps):
    return lib().orc_kl_bound(k, delta, eps)


def particle_to_matrix(state6):
    p = make_particles([state6])
    m = np.zeros(12, dtype=np.float32)
    lib().orc_particle_to_matrix(_p(p), _p(m))
    return m.reshape(3, 4)


def matrix_to_particle(m34):
    m = np.ascontiguousarray(m34, dtype=np.float32).reshape(12)
    p = np.zeros(1, dtype=PARTICLE)
    lib().orc_matrix_to_particle(_p(m), _p(p))
    return p[0]


def result_box(model_pts, state6, z_offset=-0.005):
    """Post-processing of getResult() in viz_cb (ref: src/auto_tracking.cpp:309-316, :432-466), restated with numpy:
    transformPointCloud(reference, toEigenMatrix(result) with translation.z += z_offset) -> compute3DCentroid ->
    computeCovarianceMatrixNormalized -> SelfAdjointEigenSolver (ascending eigenvalues; eigDx.col(2) = col(0) x col(1)) ->
    transformPointCloud into the eigenframe -> getMinMax3D -> box edge lengths, centre (tfinal) and axes.
    Eigen leaves the sign of an eigenvector unspecified: the largest component of the first two axes is made positive
    (extents and the box centre do not depend on that choice)."""
    f = np.float32
    pts = as_points(model_pts)
    m = particle_to_matrix(state6).astype(f)
    m[2, 3] = f(m[2, 3] + f(z_offset))
    x, y, z = pts["x"].astype(f), pts["y"].astype(f), pts["z"].astype(f)
    p = np.stack([((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] for r in range(3)], axis=1).astype(f)
    n = len(p)
    c = (p.astype(np.float64).sum(0) / n).astype(f)
    d = (p - c).astype(f)
    cov = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            cov[i, j] = (d[:, i] * d[:, j]).astype(f).astype(np.float64).sum() / n
    cov = cov.astype(f).astype(np.float64)
    w, v = np.linalg.eigh(cov)
    for col in range(2):
        k = int(np.argmax(np.abs(v[:, col])))
        if v[k, col] < 0:
            v[:, col] = -v[:, col]
    e = v.astype(f)
    e[:, 2] = np.cross(e[:, 0], e[:, 1]).astype(f)
    t = -(e.T @ c).astype(f)
    cp = (p @ e + t).astype(f)
    mn, mx = cp.min(0), cp.max(0)
    md = (f(0.5) * (mx + mn)).astype(f)
    return {"centroid": c, "axes": e, "extent": (mx - mn).astype(f), "center": (e @ md + c).astype(f), "eigenvalues": w.astype(f), "n": n}


def distance_coherence(a, b, w=1.0):
    pa, pb = make_points([a]), make_points([b])
    return lib().orc_distance_coherence(_p(pa), _p(pb), w)


def hsv_coherence(rgba_a, rgba_b, weight=1.0, hw=1.0, sw=1.0, vw=0.0):
    return lib().orc_hsv_coherence(int(rgba_a), int(rgba_b), weight, hw, sw, vw)


def particle_sample(state6, mean, cov, z6, quat_mode=1):
    p = make_particles([state6])
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    z = np.ascontiguousarray(z6, dtype=np.float32)
    lib().orc_particle_sample(_p(p), _p(mean), _p(cov), _p(z), quat_mode)
    return p[0]


def octree_approx_nearest(pts, res, queries):
    pts = as_points(pts)
    q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, 3)
    idx = np.zeros(len(q), dtype=np.int32)
    d2 = np.zeros(len(q), dtype=np.float32)
    lib().orc_octree_approx_nearest(_p(pts), len(pts), res, _p(q), len(q), _p(idx), _p(d2))
    return idx, d2


def change_detector_sequence(clouds, resolution, min_points):
    """OctreePointCloudChangeDetector fed one cloud after the other (setInputCloud, addPointsFromInputCloud,
    getPointIndicesFromNewVoxels(min_points), switchBuffers): number of point indices reported per cloud."""
    sizes = np.array([len(c) for c in clouds], dtype=np.int32)
    allp = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(c, dtype=POINT) for c in clouds])) if len(clouds) else np.zeros(0, dtype=POINT)
    found = np.zeros(len(clouds), dtype=np.int32)
    lib().orc_change_detector_sequence(_p(allp), _p(sizes), len(clouds), float(resolution), int(min_points), _p(found))
    return found


# ------------------------------------------------------------------ tracker
class Tracker:
    """CPU oracle tracker; method names follow the PCL surface (ref: src/auto_tracking.cpp:201-254)."""

    def __init__(self, kld=True):
        self._h = lib().orc_tracker_create(1 if kld else 0)
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tracker_destroy(self._h)
            self._h = None

    def set_i(self, key, v):
        assert lib().orc_set_i(self._h, key, int(v)) == 0

    def set_d(self, key, v):
        assert lib().orc_set_d(self._h, key, float(v)) == 0

    def set_vec6(self, key, v):
        a = np.ascontiguousarray(v, dtype=np.float64)
        assert len(a) == 6 and lib().orc_set_vec6(self._h, key, _p(a)) == 0

    def set_trans(self, m34):
        m = np.ascontiguousarray(m34, dtype=np.float32).reshape(-1)[:12].copy()
        lib().orc_set_trans(self._h, _p(m))

    def set_reference(self, pts):
        pts = as_points(pts)
        lib().orc_set_reference(self._h, _p(pts), len(pts))

    def set_input(self, pts):
        pts = as_points(pts)
        lib().orc_set_input(self._h, _p(pts), len(pts))

    def alias_table(self):
        """genAliasTable (Walker) of the current particle set: (a[int32], q[float64])."""
        n = lib().orc_get_particles(self._h, None, 0)
        a, q = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.float64)
        lib().orc_alias_table(self._h, _p(a), _p(q))
        return a, q

    def reset(self):
        """ParticleFilterTracker::resetTracking: clears the particle set, nothing else."""
        lib().orc_reset_tracking(self._h)

    def set_particles(self, parts):
        parts = np.ascontiguousarray(parts, dtype=PARTICLE)
        lib().orc_set_particles(self._h, _p(parts), len(parts))

    def get_particles(self):
        n = lib().orc_get_particles(self._h, None, 0)
        out = np.zeros(n, dtype=PARTICLE)
        lib().orc_get_particles(self._h, _p(out), n)
        return out

    def get_result(self):
        out = np.zeros(1, dtype=PARTICLE)
        lib().orc_get_result(self._h, _p(out))
        return out[0]

    def set_result(self, p):
        a = np.ascontiguousarray(p, dtype=PARTICLE).reshape(1)
        lib().orc_set_result(self._h, _p(a))

    def get_motion(self):
        out = np.zeros(1, dtype=PARTICLE)
        lib().orc_get_motion(self._h, _p(out))
        return out[0]

    def set_motion(self, p):
        a = np.ascontiguousarray(p, dtype=PARTICLE).reshape(1)
        lib().orc_set_motion(self._h, _p(a))

    def set_changed(self, c):
        lib().orc_set_changed(self._h, 1 if c else 0)

    def changed(self):
        return bool(lib().orc_get_changed(self._h))

    def set_change_detector(self, on, interval=10, min_points=10, resolution=0.01):
        """setUseChangeDetector / setIntervalOfChangeDetection / setMinPointsOfChangeDetection /
        setResolutionOfChangeDetection (PCL defaults; off in the reference)."""
        self.set_i(USE_CHANGE_DETECTOR, 1 if on else 0)
        self.set_i(CD_INTERVAL, int(interval))
        self.set_i(CD_MIN_POINTS, int(min_points))
        self.set_d(CD_RESOLUTION, float(resolution))

    def change_detector_info(self):
        out = np.zeros(4, dtype=np.int32)
        lib().orc_change_detector_info(self._h, _p(out))
        return {"counter": int(out[0]), "tests": int(out[1]), "last_found": int(out[2]), "changed": bool(out[3])}

    def inject_draws(self, usel, normals6, umotion):
        usel = np.ascontiguousarray(usel, dtype=np.float32)
        slots, stride = usel.shape
        normals6 = np.ascontiguousarray(normals6, dtype=np.float32).reshape(slots, stride, 6)
        umotion = np.ascontiguousarray(umotion, dtype=np.float32).reshape(slots, stride)
        lib().orc_inject_draws(self._h, _p(usel), _p(normals6), _p(umotion), slots, stride)

    def init_particles(self):
        lib().orc_init_particles(self._h)

    def resample(self, slot=0):
        lib().orc_resample(self._h, slot)

    def weight(self, keep_nn=False):
        lib().orc_weight(self._h, 1 if keep_nn else 0)

    def normalize(self):
        lib().orc_normalize(self._h)

    def update(self):
        lib().orc_update(self._h)

    def compute(self):
        lib().orc_compute(self._h)

    def aabb(self):
        a = np.zeros(6, dtype=np.float32)
        lib().orc_get_aabb(self._h, _p(a))
        return a

    def local_aabb(self):
        a = np.zeros(6, dtype=np.float32)
        lib().orc_get_local_aabb(self._h, _p(a))
        return a

    def set_crop_box(self, box6):
        if box6 is None:
            lib().orc_set_crop_box(self._h, None)
        else:
            a = np.ascontiguousarray(box6, dtype=np.float32)
            lib().orc_set_crop_box(self._h, _p(a))

    def cropped(self):
        n = lib().orc_get_cropped(self._h, None, None, 0)
        idx = np.zeros(n, dtype=np.int32)
        pts = np.zeros(n, dtype=POINT)
        lib().orc_get_cropped(self._h, _p(idx), _p(pts), n)
        return idx, pts

    def nn(self, particle, m):
        idx = np.zeros(m, dtype=np.int32)
        d2 = np.zeros(m, dtype=np.float32)
        r = lib().orc_get_nn(self._h, particle, _p(idx), _p(d2), m)
        assert r == m, (r, m)
        return idx, d2

    def raw_weights(self):
        n = lib().orc_get_raw_weights(self._h, None, 0)
        w = np.zeros(n, dtype=np.float32)
        lib().orc_get_raw_weights(self._h, _p(w), n)
        return w

    def ancestors(self):
        n = lib().orc_get_ancestors(self._h, None, 0)
        a = np.zeros(n, dtype=np.int32)
        lib().orc_get_ancestors(self._h, _p(a), n)
        return a

    def fit_ratio(self):
        return lib().orc_get_fit_ratio(self._h)

    def stage_seconds(self, reset=False):
        s = np.zeros(8, dtype=np.float64)
        lib().orc_get_stage_seconds(self._h, _p(s), 1 if reset else 0)
        return dict(zip(("transform", "crop", "index", "coherence", "normalize", "resample", "update", "total"), s.tolist()))


def configure_like_reference(t, particle_num=400, max_particle_num=500, use_hsv=True, nn_mode=NN_EXACT_BRUTE,
                             iteration_num=2, threads=0):
    """Apply the knob values of ref: src/auto_tracking.cpp:187-253."""
    step = [0.015 * 0.015] * 6
    for k in (3, 4, 5):
        step[k] *= 40.0
    t.set_i(THREADS, threads)
    t.set_i(MAX_PARTICLE_NUM, max_particle_num)
    t.set_d(DELTA, 0.99)
    t.set_d(EPSILON, 0.2)
    t.set_vec6(BIN_SIZE, [0.1] * 6)
    t.set_trans(np.eye(4, dtype=np.float32)[:3])
    t.set_vec6(STEP_COV, step)
    t.set_vec6(INIT_COV, [0.00001] * 6)
    t.set_vec6(INIT_MEAN, [0.0] * 6)
    t.set_i(ITERATION_NUM, iteration_num)
    t.set_i(PARTICLE_NUM, particle_num)
    t.set_i(NN_MODE, nn_mode)
    t.set_i(USE_DISTANCE, 1)
    t.set_i(USE_HSV, 1 if use_hsv else 0)
    t.set_d(HSV_WEIGHT, 0.1)
    t.set_d(MAX_DIST, 0.1)
    return t
