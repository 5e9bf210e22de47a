"""Deterministic synthetic LiDAR clouds.

A ring-LiDAR model: ``n_beams`` elevations x ``n_az`` azimuth steps, ground
plane plus a piecewise-constant ring of 64 obstacle sectors, range noise,
per-sweep ego shift.  Points come out sweep-major, beam-major, azimuth-minor
(the un-shuffled val-mode order).  All arithmetic in float64, cast to float32
at the end; RNG draws happen in exactly the order written so that the point
and voxel counts quoted in BASELINE.md reproduce.

Shapes follow the reference's loaders: nuScenes rows are
``x, y, z, intensity, dt`` (det3d/datasets/pipelines/loading.py:56,271-298),
Waymo rows ``x, y, z, tanh(intensity), elongation[, dt]`` (loading.py:106-115,328).
"""
import numpy as np

# Grids from the reference's configs.
GRIDS = {
    # configs/nusc/pp/polarstream/polarstream_det_n_seg_1_sector.py:10-19,63-73
    "NUSC-PILLAR": dict(range=[0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0],
                        voxel_size=[0.098, 0.0123, 8.0], max_points=20, max_voxels=60000),
    # configs/nusc/voxelnet/voxelnet_det_cylinder_singlehead.py:8-18,68-74
    "NUSC-CYL": dict(range=[0.3, -3.1488, -5.0, 50.476, 3.1488, 3.0],
                     voxel_size=[0.049, 0.00615, 0.2], max_points=30, max_voxels=180000),
    # configs/waymo/voxelnet/waymo_partner_36epoch.py:10-21,34,103-108
    "WAYMO-PARTNER": dict(range=[0.3, -3.14368, -2.0, 75.18, 3.14368, 4.0],
                          voxel_size=[0.065, 0.00307, 0.15], max_points=5, max_voxels=150000),
}


def lidar_sweep(rng, n_beams, n_az, elev_lo, elev_hi, h, max_r, dropout):
    el = np.deg2rad(np.linspace(elev_lo, elev_hi, n_beams))[:, None]
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)[None, :] + rng.uniform(0, 2 * np.pi / n_az)
    seg_r = rng.uniform(4, max_r, 64)
    seg_h = rng.uniform(0.5, 6, 64)
    seg = ((az + np.pi) / (2 * np.pi) * 64).astype(np.int64) % 64
    r_obs = np.broadcast_to(seg_r[seg], (n_beams, n_az))
    h_obs = np.broadcast_to(seg_h[seg], (n_beams, n_az))
    tan_el = np.tan(el)
    with np.errstate(divide="ignore"):
        r_gnd = np.where(el < 0, h / np.tan(-el), np.inf)
    r_gnd = np.broadcast_to(r_gnd, (n_beams, n_az))
    hit_h = h + r_obs * tan_el
    hit = (hit_h >= 0) & (hit_h <= h_obs) & (r_obs < r_gnd)
    r = np.where(hit, r_obs, r_gnd)
    r = np.where(np.isfinite(r), r, rng.uniform(20, max_r, r.shape))
    r = np.minimum(r, max_r) * (1 + rng.normal(0, 0.002, r.shape))
    x = r * np.cos(az)
    y = r * np.sin(az)
    z = r * tan_el
    p = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    if dropout > 0:
        keep = rng.random(p.shape[0]) > dropout
        p = p[keep]
    return p


def nusc_frame(seed, nsweeps=10):
    """nuScenes-shaped 10-sweep frame -> float32 [N~290k, 5]."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(nsweeps):
        p = lidar_sweep(rng, 32, 1085, -30.67, 10.67, 1.84, 100.0, 0.15)
        p[:, 0] += 0.25 * s
        close = (np.abs(p[:, 0]) < 1.0) & (np.abs(p[:, 1]) < 1.0)   # loading.py:61-70
        p = p[~close]
        inten = rng.uniform(0, 255, (p.shape[0], 1))
        dt = np.full((p.shape[0], 1), 0.05 * s)
        out.append(np.hstack([p, inten, dt]))
    return np.concatenate(out, axis=0).astype(np.float32)


def waymo_frame(seed, nsweeps=1, time_column=True):
    """Waymo-shaped frame -> float32 [169600*nsweeps, 5|6]."""
    rng = np.random.default_rng(seed)
    out = []
    for s in range(nsweeps):
        p = lidar_sweep(rng, 64, 2650, -17.6, 2.4, 2.18, 75.0, 0.0)
        p[:, 2] += 2.18
        p[:, 0] += 0.5 * s
        cols = [p, np.tanh(rng.uniform(0, 3, (p.shape[0], 1))), rng.uniform(0, 1.5, (p.shape[0], 1))]
        if time_column:
            cols.append(np.full((p.shape[0], 1), 0.1 * s))
        out.append(np.hstack(cols))
    return np.concatenate(out, axis=0).astype(np.float32)


def frame_seed(config_id, frame):
    """Seed convention of SURVEY.md section 8d: 1000 * config + frame."""
    return 1000 * int(config_id) + int(frame)


def make_batch(kind, config_id, n_frames, first_frame=0, shuffle=False, **kw):
    """List of Cartesian float32 frames for one batch."""
    fn = {"nusc": nusc_frame, "waymo": waymo_frame}[kind]
    frames = []
    for f in range(first_frame, first_frame + n_frames):
        p = fn(frame_seed(config_id, f), **kw)
        if shuffle:   # train-mode shuffle_points=True variant
            np.random.default_rng(frame_seed(config_id, f) + 500).shuffle(p, axis=0)
        frames.append(p)
    return frames
