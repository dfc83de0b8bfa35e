"""Generate tests/golden/refgen_golden.npz with the REFERENCE's own reference-generation code
(needs /root/reference; the GPU box does not have it, hence the committed fixture).

Executed from the reference tree:
  cmd_pc/scripts/traj_gen/polym_optimizer.py   PolymOptimizer.get_coeff  (min-snap xyz, min-accel yaw)
  ndp_nmpc/scripts/pt_pub/base_pt_publisher.py BasePtPublisher.get_traj_pt (piecewise polynomial evaluation,
                                                hover after the end)
  ndp_nmpc/scripts/pt_pub/pt_publisher.py      diff_flatness, NMPCRefPublisher.traj_full_pt_2_x_u
ROS is not installed here, so `rospy`, the message classes and `tf_conversions` are replaced by minimal
stand-ins below (plain attribute containers; a Time type with to_sec()).  The only arithmetic inside a
stand-in is tf.transformations.quaternion_from_matrix, restated from the ROS `tf` package's published
algorithm [EXT]; everything else is the reference's code.  Time allocation t = d / v_mean restates
cmd_pc/scripts/traj_gen/traj_generator.py:55-64 (that module imports rospy message types at import time).
"""
import math
import os
import sys
import types

import numpy as np
import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
REF = "/root/reference"


# ---------------- stand-ins for the ROS python modules ----------------
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _vec():
    return _NS(x=0.0, y=0.0, z=0.0)


class Time:
    def __init__(self, sec=0.0):
        self.sec = float(sec)

    def __sub__(self, o):
        return Duration(self.sec - o.sec)

    def __add__(self, d):
        return Time(self.sec + d.sec)

    @staticmethod
    def now():
        return Time(0.0)


class Duration:
    def __init__(self, sec=0.0):
        self.sec = float(sec)

    def to_sec(self):
        return self.sec

    @staticmethod
    def from_sec(s):
        return Duration(s)


def quaternion_from_matrix(matrix):
    """ROS tf.transformations.quaternion_from_matrix (x, y, z, w) [EXT]."""
    q = np.empty((4,), dtype=np.float64)
    M = np.array(matrix, dtype=np.float64)[:4, :4]
    t = np.trace(M)
    if t > M[3, 3]:
        q[3] = t
        q[2] = M[1, 0] - M[0, 1]
        q[1] = M[0, 2] - M[2, 0]
        q[0] = M[2, 1] - M[1, 2]
    else:
        i, j, k = 0, 1, 2
        if M[1, 1] > M[0, 0]:
            i, j, k = 1, 2, 0
        if M[2, 2] > M[i, i]:
            i, j, k = 2, 0, 1
        t = M[i, i] - (M[j, j] + M[k, k]) + M[3, 3]
        q[i] = t
        q[j] = M[i, j] + M[j, i]
        q[k] = M[k, i] + M[i, k]
        q[3] = M[k, j] - M[j, k]
    q *= 0.5 / math.sqrt(t * M[3, 3])
    return q


def install_stubs():
    rospy = types.ModuleType("rospy")
    rospy.Time, rospy.Duration = Time, Duration
    sys.modules["rospy"] = rospy
    tfc = types.ModuleType("tf_conversions")
    tfc.transformations = _NS(quaternion_from_matrix=quaternion_from_matrix, euler_from_quaternion=None)
    sys.modules["tf_conversions"] = tfc
    for name, classes in (("geometry_msgs.msg", {"Point": _vec}), ("nav_msgs.msg", {"Odometry": _NS})):
        pkg = types.ModuleType(name.split(".")[0]); sys.modules[name.split(".")[0]] = pkg
        m = types.ModuleType(name); sys.modules[name] = m
        for k, v in classes.items():
            setattr(m, k, v)

    class TrajPt:  # msg/TrajPt.msg
        def __init__(self):
            self.position, self.velocity, self.accel, self.jerk = _vec(), _vec(), _vec(), _vec()
            self.yaw, self.yaw_dot = 0.0, 0.0

    class TrajFullStatePt:  # msg/TrajFullStatePt.msg
        def __init__(self):
            self.pose = _NS(position=_vec(), orientation=_NS(x=0.0, y=0.0, z=0.0, w=0.0))
            self.twist = _NS(linear=_vec(), angular=_vec())
            self.collective_force = 0.0

    class TrajCoefficients:  # msg/TrajCoefficients.msg
        def __init__(self):
            self.coeff_x, self.coeff_y, self.coeff_z, self.coeff_yaw = [], [], [], []
            self.traj_time_cum, self.traj_time_seg, self.final_pt = [], [], _vec()

    pkg = types.ModuleType("ndp_nmpc"); sys.modules["ndp_nmpc"] = pkg
    m = types.ModuleType("ndp_nmpc.msg"); sys.modules["ndp_nmpc.msg"] = m
    m.TrajPt, m.TrajFullStatePt, m.TrajCoefficients = TrajPt, TrajFullStatePt, TrajCoefficients
    return TrajCoefficients


def main():
    TrajCoefficients = install_stubs()
    sys.path.insert(0, os.path.join(REF, "ndp_nmpc", "scripts"))
    from pt_pub.pt_publisher import NMPCRefPublisher  # reference module

    import importlib.util
    spec = importlib.util.spec_from_file_location("cmd_pc_polym", os.path.join(REF, "cmd_pc/scripts/traj_gen/polym_optimizer.py"))
    cpo = importlib.util.module_from_spec(spec); spec.loader.exec_module(cpo)  # reference module (planner side)

    out = {}
    rng = np.random.default_rng(0)
    for name in ("eight_high_dyn", "eight_low", "eight_low_diff_h"):
        path = yaml.safe_load(open(os.path.join(REF, "cmd_pc/path_config", name + ".yaml")))["path"]
        xyz = np.array([p["pos"] for p in path], dtype=np.float64).T
        yaw = np.radians(np.array([p["yaw"] for p in path], dtype=np.float64))  # load_path.py:18
        spd = np.array([p["vel"] for p in path], dtype=np.float64)
        dist = xyz[:, 1:] - xyz[:, :-1]
        t_seg = np.sqrt((dist**2).sum(0)) / ((spd[:-1] + spd[1:]) / 2)  # traj_generator.py:55-64
        t_cum = np.insert(np.cumsum(t_seg), 0, 0.0)
        tc = TrajCoefficients()
        tc.coeff_x = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[0])).tolist()
        tc.coeff_y = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[1])).tolist()
        tc.coeff_z = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[2])).tolist()
        tc.coeff_yaw = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.ACCEL).get_coeff(yaw)).tolist()
        tc.traj_time_cum, tc.traj_time_seg = t_cum.tolist(), t_seg.tolist()
        tc.final_pt.x, tc.final_pt.y, tc.final_pt.z = xyz[0, -1], xyz[1, -1], xyz[2, -1]
        pub = NMPCRefPublisher()
        pub.traj_coeff, pub.start_ros_t = tc, Time(0.0)
        # sample times: random inside, exactly on knots, and beyond the end (hover branch)
        ts = np.concatenate([rng.uniform(0, t_cum[-1], 60), t_cum[:-1], [t_cum[-1], t_cum[-1] + 0.5, t_cum[-1] - 1e-9]])
        X, U = [], []
        for t in ts:
            x, u = pub.traj_full_pt_2_x_u(pub.get_traj_full_state_pt(Time(t)))
            X.append(x); U.append(u)
        out[name + "_wpts"], out[name + "_yaw"], out[name + "_speed"] = xyz, yaw, spd
        out[name + "_t_cum"] = t_cum
        out[name + "_cx"], out[name + "_cy"], out[name + "_cz"] = np.array(tc.coeff_x), np.array(tc.coeff_y), np.array(tc.coeff_z)
        out[name + "_cyaw"] = np.array(tc.coeff_yaw)
        out[name + "_t"], out[name + "_x"], out[name + "_u"] = ts, np.array(X), np.array(U)
        print(name, "T =", t_cum[-1], "segments", len(t_seg), "x(0.7T) =", X[0][:3])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refgen_golden.npz"), **out)


if __name__ == "__main__":
    main()
