"""TEST INFRASTRUCTURE ONLY — never imported by remixfusion_b200/.

NumPy restatement of the tracker kernels of the reference (model/ROtracker.py:141-400), one array expression per
source line, vectorised over pixels / (candidate, pixel) pairs:

  * vertex_map   compute_vertex  :273-344 — given the per-ROW random sample (the reference draws it with curand XORWOW,
                 subsequence = row index; the generator itself is NVIDIA's and is not restated: golden vectors carry the
                 samples the literal kernel drew)
  * normal_map   compute_normal  :346-400
  * fitness      compute_tsdf_value :144-271 (+ host evaluate_tsdf :536-604)
  * cal_transform  the host loop of :606-714, statement by statement

PARITY: pinned by tests/golden/track_golden.npz = outputs of the literal reference kernels (oracle/_ref/ref_tracker.cubin)
on a B200 (tests/golden/make_track_golden.py).  fp32 throughout; fused multiply-adds are emulated where the compiled
reference kernel has them (fma32: exact product and sum in float64, one rounding to float32 — double rounding can
differ from a hardware FMA by one ulp in rare cases, hence the tolerances in tests/test_track_oracle.py)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def vertex_map(depth, K, cut_dist, trunc, row_sample, sample_range):
    """depth [H,W] fp32, row_sample [H] fp32 -> depth_vertex [H,W,4]."""
    H, W = depth.shape
    K = np.asarray(K, f32).reshape(-1)
    d = depth.astype(f32).copy()
    d[d > f32(cut_dist)] = 0                                               # :292-294
    sample = row_sample.astype(f32)[:, None] * np.ones((1, W), f32)
    z_val = sample * f32(trunc)
    gt = -sample                                                           # :326-334
    gt = np.where(z_val < f32(-1) * f32(trunc), f32(1), gt)
    gt = np.where(z_val > f32(1) * f32(trunc), f32(1), gt)
    c_z = d + z_val                                                        # :337-339
    pj = np.arange(W, dtype=f32)[None, :]; pi = np.arange(H, dtype=f32)[:, None]
    c_x = ((pj - K[2]) * c_z) / K[0]
    c_y = ((pi - K[5]) * c_z) / K[4]
    out = np.stack([c_x, c_y, c_z, gt.astype(f32)], -1).astype(f32)
    out[d <= 0] = 0                                                        # :296-303
    return out


def normal_map(vertex):
    """vertex [H,W,4] -> normal [H,W,3]; border pixels stay zero (:353-355)."""
    H, W, _ = vertex.shape
    n = np.zeros((H, W, 3), f32)
    c = vertex[1:-1, 1:-1]; l = vertex[1:-1, :-2]; r = vertex[1:-1, 2:]; u = vertex[:-2, 1:-1]; d = vertex[2:, 1:-1]
    ok = (c[..., 2] != 0) & (l[..., 2] != 0) & (r[..., 2] != 0) & (u[..., 2] != 0) & (d[..., 2] != 0)
    hor = l[..., :3] - r[..., :3]; ver = u[..., :3] - d[..., :3]
    # :379-385 as compiled: a*b + c*d -> fma(a, b, c*d)
    nx = fma32(hor[..., 1], ver[..., 2], (-hor[..., 2]) * ver[..., 1])
    ny = fma32(hor[..., 2], ver[..., 0], -(hor[..., 0] * ver[..., 2]))
    nz = fma32(hor[..., 0], ver[..., 1], (-hor[..., 1]) * ver[..., 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        lens = np.sqrt(fma32(nz, nz, fma32(ny, ny, nx * nx)))
        nx, ny, nz = nx / lens, ny / lens, nz / lens
    flip = nz > 0
    nx = np.where(flip, -nx, nx); ny = np.where(flip, -ny, ny); nz = np.where(flip, -nz, nz)
    inner = np.stack([nx, ny, nz], -1).astype(f32)
    inner[~ok] = 0
    n[1:-1, 1:-1] = inner
    return n


def fitness(tsdf, vol_dim, vol_origin, voxel, vertex, normal, K, R, T, cand, search_size, level, level_index):
    """tsdf flat fp32 (index z + y*dz + x*dy*dz); vertex [H,W,4]; normal [H,W,3]; cand [n,6] -> (value [n], count [n])."""
    H, W, _ = vertex.shape
    K = np.asarray(K, f32).reshape(-1); R = np.asarray(R, f32).reshape(-1); T = np.asarray(T, f32).reshape(-1)
    ss = np.asarray(search_size, f32).reshape(-1); cand = np.asarray(cand, f32)
    dx, dy, dz = (int(v) for v in vol_dim)
    ox, oy, oz = (f32(int(v)) for v in vol_origin)                         # :163-165 origin truncated to int
    ph, pw = int(H / level), int(W / level)
    pi = (np.arange(ph) * level + level_index)[:, None] * np.ones((1, pw), int)
    pj = (np.arange(pw) * level + level_index)[None, :] * np.ones((ph, 1), int)
    ok = (pi <= ph * level - 1) & (pj <= pw * level - 1)
    pi, pj = pi[ok], pj[ok]
    nrm = normal[pi, pj]; v = vertex[pi, pj]
    keep = ~((nrm == 0).all(-1)) & ~((v[:, :3] == 0).all(-1))              # :186, :201
    v = v[keep]
    x, y, z, gt = v[:, 0], v[:, 1], v[:, 2], v[:, 3]
    gx = fma32(z, R[2], fma32(x, R[0], y * R[1]))                           # :211-213 (as compiled)
    gy = fma32(z, R[5], fma32(x, R[3], y * R[4]))
    gz = fma32(z, R[8], fma32(x, R[6], y * R[7]))
    q1 = (cand[:, 3] * ss[3])[:, None]; q2 = (cand[:, 4] * ss[4])[:, None]; q3 = (cand[:, 5] * ss[5])[:, None]   # :219-221
    q0 = np.sqrt(fma32(-q3, q3, fma32(-q2, q2, fma32(-q1, q1, f32(1)))))    # :222
    gx, gy, gz, gt = gx[None], gy[None], gz[None], gt[None]
    q_z = fma32(gz, q0, fma32(gy, q1, -(gx * q2)))                          # :224-227
    S = fma32(gz, q3, fma32(gx, q1, gy * q2))
    q_y = fma32(-gz, q1, fma32(gx, q3, gy * q0))
    q_x = fma32(gz, q2, fma32(gx, q0, -(gy * q3)))
    c0, c1, c2 = cand[:, 0:1], cand[:, 1:2], cand[:, 2:3]
    X = fma32(c0, ss[0], fma32(-q3, q_y, fma32(q2, q_z, fma32(q1, S, q_x * q0)))) + T[0]      # :229-231
    Y = fma32(c1, ss[1], fma32(q3, q_x, fma32(q2, S, fma32(q_y, q0, -(q1 * q_z))))) + T[1]
    Z = fma32(c2, ss[2], fma32(q3, S, fma32(-q2, q_x, fma32(q_z, q0, q1 * q_y)))) + T[2]
    vx, vy, vz = X - T[0], Y - T[1], Z - T[2]                               # :233-235
    cam_x = fma32(R[6], vz, fma32(R[0], vx, R[3] * vy))                     # :237-239
    cam_y = fma32(R[7], vz, fma32(R[1], vx, R[4] * vy))
    cam_z = fma32(R[8], vz, fma32(R[2], vx, R[5] * vy))
    with np.errstate(invalid="ignore", divide="ignore"):
        px = np.trunc((K[2] + (cam_x * K[0]) / cam_z) + f32(0.5))           # :241-242
        py = np.trunc((K[5] + (cam_y * K[4]) / cam_z) + f32(0.5))
        hit = (px >= 0) & (py >= 0) & (px < W) & (py < H) & (cam_z >= 0)    # :245
        rnd = lambda a: np.sign(a) * np.floor(np.abs(a) + f32(0.5))         # roundf: half away from zero
        vxi = rnd((X - ox) / f32(voxel)); vyi = rnd((Y - oy) / f32(voxel)); vzi = rnd((Z - oz) / f32(voxel))   # :246-248
    inside = (vxi >= 1) & (vxi < dx - 1) & (vyi >= 1) & (vyi < dy - 1) & (vzi >= 1) & (vzi < dz - 1)            # :250
    hit &= inside
    idx = (np.where(hit, vzi, 0) + np.where(hit, vyi, 0) * dz + np.where(hit, vxi, 0) * dy * dz).astype(np.int64)   # :254
    add = np.abs(tsdf[idx] - gt).astype(f32)                                # :261
    add = np.where(hit, add, f32(0))
    value = add.astype(np.float64).sum(1).astype(f32)                       # the reference sums with fp32 atomics in arbitrary order
    count = hit.sum(1).astype(f32)
    return value, count


def cal_transform(search_value, transform_candidate, search_size, count_search_max):
    """model/ROtracker.py:606-714, literally (Python floats accumulate float32 products, as under NumPy 1.x)."""
    search_value = np.asarray(search_value, f32); cand = np.asarray(transform_candidate, f32); ss = np.asarray(search_size, f32)
    mean_transform = np.zeros(7, f32)
    origin_tsdf = search_value[0]
    sum_tx = sum_ty = sum_tz = sum_qw = sum_qx = sum_qy = sum_qz = sum_weight = sum_tsdf = 0.0
    count_search = 0
    for j in range(1, len(search_value)):
        if search_value[j] < origin_tsdf:
            tx, ty, tz, qx, qy, qz = (cand[j][k] for k in range(6))
            cur_fit = search_value[j]
            weight = origin_tsdf - cur_fit
            sum_tx += float(tx * weight); sum_ty += float(ty * weight); sum_tz += float(tz * weight)
            sum_qx += float(qx * weight); sum_qy += float(qy * weight); sum_qz += float(qz * weight)
            qx = qx * ss[3]; qy = qy * ss[4]; qz = qz * ss[5]
            qw = np.sqrt(f32(1) - qx * qx - qy * qy - qz * qz)
            sum_qw += float(qw * weight); sum_weight += float(weight); sum_tsdf += float(cur_fit * weight)
            count_search += 1
            if count_search == count_search_max:
                break
    if count_search <= 0:
        return False, float(origin_tsdf), mean_transform
    mean_tsdf = sum_tsdf / sum_weight
    mean_transform[0] = (sum_tx / sum_weight) * float(ss[0])
    mean_transform[1] = (sum_ty / sum_weight) * float(ss[1])
    mean_transform[2] = (sum_tz / sum_weight) * float(ss[2])
    qww = sum_qw / sum_weight
    qxx = (sum_qx / sum_weight) * float(ss[3]); qyy = (sum_qy / sum_weight) * float(ss[4]); qzz = (sum_qz / sum_weight) * float(ss[5])
    lens = 1 / np.sqrt(qww * qww + qxx * qxx + qyy * qyy + qzz * qzz)
    mean_transform[3] = qww * lens; mean_transform[4] = qxx * lens; mean_transform[5] = qyy * lens; mean_transform[6] = qzz * lens
    return True, float(mean_tsdf), mean_transform


def update_PST(search_size, tsdf, mean_transform, min_scale=1e-3, scale=0.09):
    """model/ROtracker.py:495-531, statement by statement; writes into `search_size` (float32[6]) in place.
    Types as under the reference's numpy 1.21.6 (requirements.txt:6): a float32 scalar combined with a Python float or int gives
    float64 there (before NEP 50), so every scalar below is float64 and only the stores into the float32 array narrow — made
    explicit with float() so that the restatement does not depend on the NumPy version that runs it."""
    tsdf = float(tsdf)
    s_tx = abs(float(mean_transform[0])) + min_scale
    s_ty = abs(float(mean_transform[1])) + min_scale
    s_tz = abs(float(mean_transform[2])) + min_scale
    s_qx = abs(float(mean_transform[4])) + min_scale
    s_qy = abs(float(mean_transform[5])) + min_scale
    s_qz = abs(float(mean_transform[6])) + min_scale
    trans_norm = np.sqrt(s_tx**2 + s_ty**2 + s_tz**2 + s_qx**2 + s_qy**2 + s_qz**2)
    normal_tx = s_tx / trans_norm; normal_ty = s_ty / trans_norm; normal_tz = s_tz / trans_norm
    normal_qx = s_qx / trans_norm; normal_qy = s_qy / trans_norm; normal_qz = s_qz / trans_norm
    search_size[3] = scale * tsdf * normal_qx + min_scale
    search_size[4] = scale * tsdf * normal_qy + min_scale
    search_size[5] = scale * tsdf * normal_qz + min_scale
    search_size[0] = scale * tsdf * normal_tx + min_scale
    search_size[1] = scale * tsdf * normal_ty + min_scale
    search_size[2] = scale * tsdf * normal_tz + min_scale


def random_optimization(tr, cur_id, cam_pose, depth_im, cam_intr, beta=0.9, inherit=False, seed_num=None):
    """The search loop of model/ROtracker.py:716-836, statement by statement, driving a tracker object `tr` that offers the
    reference's step methods (init_depth_vertex / init_normal / evaluate_tsdf / cal_transform / init_searchsize) and attributes
    (tiff_index, depth_level, PST_size, ALL_PST, particle_iter_lens, fix_level_index, scaling_coefficient, iterative_scale).
    Returns (pose 4x4, list of per-iteration success flags)."""
    def get_PST(tiff_index):                                                             # :467-493
        PST_class = tiff_index // 20
        PST_class_num = tiff_index - PST_class * 20
        PST_class_index = PST_class_num // 3
        return tr.ALL_PST[PST_class][PST_class_index, ...]
    tr.current_global_R = cam_pose[:3, :3].copy()
    tr.current_global_T = cam_pose[:3, 3].copy()
    if inherit is True and tr.previous_frame_success:
        tr.search_size = tr.initialize_search_size
    else:
        tr.init_searchsize()
    tr.init_depth_vertex(depth_im, cam_intr, seed_num=seed_num)
    tr.init_normal()
    previous_success = False
    success = False
    count_particle = 0
    level_index = 5
    flags = []
    for i in range(tr.particle_iter_lens):
        if not success:
            count_particle = 0
        PST_class = count_particle % 3
        tr.transform_candidate = get_PST(tr.tiff_index[count_particle])
        level = tr.depth_level[count_particle]
        search_value, sv, sc = tr.evaluate_tsdf(cur_id, level, tr.PST_size[PST_class], cam_intr, level_index)
        success, min_tsdf, mean_transform = tr.cal_transform(search_value)
        flags.append(bool(success))
        current_T_incremental = mean_transform[:3]
        qw = mean_transform[3]; qx = mean_transform[4]; qy = mean_transform[5]; qz = mean_transform[6]
        if success:
            if count_particle < 19:
                count_particle += 1
            # float32 products and sums; `2 *` and `1 -` promote to float64 under numpy 1.21 (see update_PST), narrowed by dtype=
            current_R_incremental = np.array([
                [1 - 2 * float(qy*qy + qz*qz), 2 * float(qx*qy - qz*qw),     2 * float(qx*qz + qy*qw)],
                [2 * float(qx*qy + qz*qw),     1 - 2 * float(qx*qx + qz*qz), 2 * float(qy*qz - qx*qw)],
                [2 * float(qx*qz - qy*qw),     2 * float(qy*qz + qx*qw),     1 - 2 * float(qx*qx + qy*qy)]
            ], dtype=np.float32)
            tr.current_global_T += current_T_incremental
            # np.matmul of two float32 3x3 matrices: written out (float32 products summed left to right) so that the result does
            # not depend on the BLAS behind NumPy; any summation order differs from this one by at most an ulp per entry
            A, B = current_R_incremental, np.asarray(tr.current_global_R, dtype=np.float32)
            tr.current_global_R = np.array([[(A[r, 0] * B[0, c] + A[r, 1] * B[1, c]) + A[r, 2] * B[2, c] for c in range(3)] for r in range(3)],
                                           dtype=np.float32)
        if tr.fix_level_index:
            level_index = 1
        else:
            level_index += 5
        level_index = level_index % (tr.depth_level[count_particle])
        update_PST(tr.search_size, min_tsdf, mean_transform, scale=tr.scaling_coefficient)
        if previous_success and success:
            for k in range(6):
                tr.search_size[k] = beta * float(tr.search_size[k]) + (1 - beta) * float(tr.previous_search_size[k])
        elif success:
            if tr.iterative_scale:
                previous_success = True
            for k in range(6):
                tr.previous_search_size[k] = tr.search_size[k]
        if not success:
            previous_success = False
        if i == 0:
            if success:
                tr.initialize_search_size = tr.search_size
                tr.previous_frame_success = True
            else:
                tr.previous_frame_success = False
    cam_pose_iter = np.eye(4, dtype=np.float32)
    cam_pose_iter[:3, :3] = tr.current_global_R
    cam_pose_iter[:3, 3] = tr.current_global_T
    return cam_pose_iter, flags
