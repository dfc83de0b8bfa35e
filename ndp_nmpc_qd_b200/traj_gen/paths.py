"""Way-point tables of the reference's figure-eight paths (data of cmd_pc/path_config/*.yaml:
pos [m], yaw [deg], vel [m/s]); loaded by the reference with load_path.py:13-22 (yaw -> radians)."""
import numpy as np


def _eight(center, ax, ay, dz, laps, vel, yaw_deg=None):
    cx, cy, cz = center
    lap = [(cx + ax, cy + ay, cz + dz), (cx + 2 * ax, cy, cz + 2 * dz), (cx + ax, cy - ay, cz + dz), (cx, cy, cz),
           (cx - ax, cy + ay, cz - dz), (cx - 2 * ax, cy, cz - 2 * dz), (cx - ax, cy - ay, cz - dz), (cx, cy, cz)]
    pts = [(cx, cy, cz)] + lap * laps
    yaw = [0.0] * len(pts) if yaw_deg is None else yaw_deg
    return dict(pos=np.array(pts, dtype=np.float64), yaw_deg=np.array(yaw, dtype=np.float64), vel=np.full(len(pts), float(vel)))


PATHS = {
    # eight_high_dyn.yaml: 17 way-points, two laps, 5 m/s
    "eight_high_dyn": _eight((1.0, 1.0, 5.0), 5.0, 5.0, 1.5, 2, 5.0),
    # eight_low.yaml / eight_middle.yaml: 9 way-points, one lap, 0.5 m/s
    "eight_low": _eight((1.0, 1.0, 0.5), 2.0, 1.0, 0.0, 1, 0.5),
    "eight_middle": _eight((1.0, -0.5, 1.5), 2.0, -1.0, 0.0, 1, 0.5),
    # eight_low_diff_h.yaml: height and yaw vary
    "eight_low_diff_h": dict(
        pos=np.array([[1, 1, .5], [3, 2, 1], [5, 1, 1], [3, 0, 1], [1, 1, .5], [-1, 2, 0], [-3, 1, 0], [-1, 0, 0], [1, 1, .5]], dtype=np.float64),
        yaw_deg=np.array([0, 10, 20, 10, 0, -10, -20, -10, 0], dtype=np.float64), vel=np.full(9, 0.5)),
}
