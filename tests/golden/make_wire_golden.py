"""Generate tests/golden/wire_golden.npz by executing the REFERENCE's own node methods in this container
(needs /root/reference; the GPU box does not have it, hence the committed fixture).

What runs from the reference tree (file:line under /root/reference):
  dop_sim/scripts/dop_qd_node.py:131-148   DopQdNode.pub_odom_callback       plant state -> nav_msgs/Odometry
  ndp_nmpc/scripts/pt_pub/pt_publisher.py:106-122  NMPCRefPublisher.odom_2_nmpc_x   Odometry -> x0
  ndp_nmpc/scripts/nmpc_node.py:273-283    ControllerNode.nmpc_u_2_att_tgt   u0 -> mavros_msgs/AttitudeTarget
  dop_sim/scripts/dop_qd_node.py:162-166   DopQdNode.sub_body_rate_cmd_cb    AttitudeTarget -> plant command
  ndp_nmpc/scripts/nmpc_node.py:116-133    ControllerNode.do_pub_ref         (xr, ur) -> PredXU
  ndp_nmpc/scripts/nmpc_follower_node.py:44-74   FollowerNode.sub_formation_ref_callback / sub_pred_callback
  ndp_nmpc/scripts/ndp_nmpc_leader_node.py:49-76 NDPLeaderNode.pub_formation_ref_callback / sub_xf_pred_callback
  ndp_nmpc/scripts/pt_pub/pt_publisher.py:57-103 NMPCRefPublisher.reset / get_nmpc_pts  (the 101-point sliding list)

ROS, acados and CasADi are not installed here: `rospy`, `tf2_ros`, `actionlib`, the message packages, `tf_conversions`,
`acados_template` and `casadi` are replaced by inert stand-ins (attribute containers, a Time type).  The node classes are
never constructed (their __init__ talks to a ROS master); their methods are called unbound on plain objects that carry
the attributes the method reads.  The only arithmetic inside a stand-in is tf's quaternion_from_matrix /
euler_from_quaternion, restated from the ROS `tf` package [EXT]; everything else is the reference's code.
"""
import math
import os
import sys
import types
from unittest import mock

import numpy as np
import torch
import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_refgen_golden as mrg  # noqa: E402  (Time / Duration / quaternion_from_matrix stand-ins)


class NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Point(NS):
    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(x=x, y=y, z=z)


class Quaternion(NS):
    def __init__(self, x=0.0, y=0.0, z=0.0, w=0.0):
        super().__init__(x=x, y=y, z=z, w=w)


class Odometry:  # nav_msgs/Odometry
    def __init__(self):
        self.header = NS(stamp=None, frame_id="")
        self.pose = NS(pose=NS(position=Point(), orientation=Quaternion()))
        self.twist = NS(twist=NS(linear=Point(), angular=Point()))


class AttitudeTarget:  # mavros_msgs/AttitudeTarget
    IGNORE_ATTITUDE = 128

    def __init__(self):
        self.type_mask, self.body_rate, self.thrust = 0, Point(), 0.0


class Float64MultiArray:
    def __init__(self):
        self.data = []


class PredXU:  # ndp_nmpc/msg/PredXU.msg
    def __init__(self):
        self.header, self.x, self.u = NS(stamp=None, frame_id=""), [], []


def euler_from_quaternion(q):
    """tf.transformations.euler_from_quaternion, axes 'sxyz', q = (x, y, z, w) [EXT]."""
    x, y, z, w = q
    return (math.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)), math.asin(max(-1.0, min(1.0, 2 * (w * y - z * x)))),
            math.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z)))


CLOCK = [0.0]


def install_stubs():
    TrajCoefficients = mrg.install_stubs()

    class Time(mrg.Time):
        @staticmethod
        def now():
            return Time(CLOCK[0])

        def __add__(self, d):
            return Time(self.sec + d.sec)

    rospy = sys.modules["rospy"]
    rospy.Time = Time
    for n in ("init_node", "get_namespace", "Subscriber", "Publisher", "Timer", "loginfo", "logwarn", "loginfo_throttle", "get_param",
              "get_name", "is_shutdown", "Rate", "spin"):
        setattr(rospy, n, mock.MagicMock())
    rospy.timer = NS(TimerEvent=object)
    rospy.ROSInterruptException = Exception
    sys.modules["tf_conversions"].transformations.euler_from_quaternion = euler_from_quaternion
    for name in ("tf2_ros", "actionlib", "acados_template", "casadi", "matplotlib", "matplotlib.pyplot", "visualization_msgs",
                 "visualization_msgs.msg", "mavros_msgs", "std_msgs"):
        sys.modules.setdefault(name, mock.MagicMock())
    mm = types.ModuleType("mavros_msgs.msg")
    mm.AttitudeTarget, mm.State, mm.ESCStatus, mm.ESCStatusItem = AttitudeTarget, NS, mock.MagicMock(), mock.MagicMock()
    sys.modules["mavros_msgs.msg"] = mm
    sm = types.ModuleType("std_msgs.msg")
    sm.Float64MultiArray, sm.ColorRGBA = Float64MultiArray, mock.MagicMock()
    sys.modules["std_msgs.msg"] = sm
    gm = sys.modules["geometry_msgs.msg"]
    gm.Point, gm.Quaternion = Point, Quaternion
    for n in ("Pose", "PoseArray", "TransformStamped", "Vector3"):
        setattr(gm, n, mock.MagicMock())
    sys.modules["nav_msgs.msg"].Odometry = Odometry
    nm = sys.modules["ndp_nmpc.msg"]
    nm.PredXU = PredXU
    for n in ("TrackTrajAction", "TrackTrajGoal", "TrackTrajResult", "TrackTrajFeedback"):
        setattr(nm, n, mock.MagicMock())
    return TrajCoefficients, Time


def traj_coefficients(TrajCoefficients, name):
    """The planner's TrajCoefficients for a path YAML (cmd_pc/scripts/traj_gen, as in make_refgen_golden.py)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("cmd_pc_polym", os.path.join(REF, "cmd_pc/scripts/traj_gen/polym_optimizer.py"))
    cpo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cpo)
    path = yaml.safe_load(open(os.path.join(REF, "cmd_pc/path_config", name + ".yaml")))["path"]
    xyz = np.array([p["pos"] for p in path], dtype=np.float64).T
    yaw = np.radians(np.array([p["yaw"] for p in path], dtype=np.float64))
    spd = np.array([p["vel"] for p in path], dtype=np.float64)
    dist = xyz[:, 1:] - xyz[:, :-1]
    t_seg = np.sqrt((dist**2).sum(0)) / ((spd[:-1] + spd[1:]) / 2)
    t_cum = np.insert(np.cumsum(t_seg), 0, 0.0)
    tc = TrajCoefficients()
    tc.coeff_x = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[0])).tolist()
    tc.coeff_y = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[1])).tolist()
    tc.coeff_z = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.SNAP).get_coeff(xyz[2])).tolist()
    tc.coeff_yaw = np.squeeze(cpo.PolymOptimizer(cpo.MinMethod.ACCEL).get_coeff(yaw)).tolist()
    tc.traj_time_cum, tc.traj_time_seg = t_cum.tolist(), t_seg.tolist()
    tc.final_pt.x, tc.final_pt.y, tc.final_pt.z = xyz[0, -1], xyz[1, -1], xyz[2, -1]
    return tc


def main():
    TrajCoefficients, Time = install_stubs()
    sys.path.insert(0, os.path.join(REF, "dop_sim", "scripts"))
    sys.path.insert(0, os.path.join(REF, "ndp_nmpc", "scripts"))
    import nmpc_node as ref_node  # reference modules
    import nmpc_follower_node as ref_follower
    import ndp_nmpc_leader_node as ref_leader
    import dop_qd_node as ref_sim
    from pt_pub import NMPCRefPublisher

    rng = np.random.default_rng(0)
    out = {}

    # ---- 1. plant state -> Odometry -> x0 ----
    n = 16
    state = rng.normal(size=(n, 35, 1))
    sim = NS(num_agent=n, ego_states=torch.from_numpy(state), mul_odom=[Odometry() for _ in range(n)],
             mul_odom_pub=[mock.MagicMock() for _ in range(n)])
    ref_sim.DopQdNode.pub_odom_callback(sim, None)
    x0 = np.stack([NMPCRefPublisher.odom_2_nmpc_x(o) for o in sim.mul_odom])
    out["odom_state"], out["odom_x0"] = state[:, :, 0], x0

    # ---- 2. u0 -> AttitudeTarget -> plant command ----
    u0 = rng.normal(size=(n, 4)) * np.array([2, 2, 1, 3]) + np.array([0, 0, 0, 9.81])
    k_thr = np.concatenate([[50.0, 53.07998220238106, 0.0, 41.6], rng.uniform(40, 60, n - 4)])
    cmd = torch.zeros((n, 4, 1), dtype=torch.float64)
    sim2 = NS(body_rate_cmd=cmd)
    for i in range(n):
        att = ref_node.ControllerNode.nmpc_u_2_att_tgt(NS(k_throttle=float(k_thr[i])), *u0[i])
        assert att.type_mask == AttitudeTarget.IGNORE_ATTITUDE
        ref_sim.DopQdNode.sub_body_rate_cmd_cb(sim2, att, i)
    out["att_u0"], out["att_k_throttle"], out["att_cmd"] = u0, k_thr, cmd.numpy()[:, :, 0]

    # ---- 3. PredXU: publish (leader) / consume with formation offset (follower) ----
    B = 5
    xr, ur = rng.normal(size=(B, 21, 10)), rng.normal(size=(B, 20, 4))
    offs = rng.normal(size=(B, 3))
    msgs, fx, fu = [], [], []
    for b in range(B):
        sent = []
        node = NS(nmpc_ctl=NS(solver=NS(N=20)), nmpc_x_ref=xr[b], nmpc_u_ref=ur[b], pub_ref_x_u=NS(publish=sent.append))
        ref_node.ControllerNode.do_pub_ref(node)
        m = sent[0]
        msgs.append(np.concatenate([np.concatenate([np.array(a.data) for a in m.x]), np.concatenate([np.array(a.data) for a in m.u])]))
        fol = NS(formation_ref=Point(*offs[b]), nmpc_x_ref=np.zeros((21, 10)), nmpc_u_ref=np.zeros((20, 4)), is_print_error=False)
        ref_follower.FollowerNode.sub_pred_callback(fol, m)
        fx.append(fol.nmpc_x_ref.copy()); fu.append(fol.nmpc_u_ref.copy())
    out["predxu_xr"], out["predxu_ur"], out["predxu_msg"], out["predxu_offset"] = xr, ur, np.stack(msgs), offs
    out["predxu_follower_xr"], out["predxu_follower_ur"] = np.stack(fx), np.stack(fu)

    # ---- 4. formation references: leader's 20 Hz switch + the followers' alpha filters ----
    leader_x = np.concatenate([np.linspace(1.0, 5.0, 40), np.linspace(5.0, -3.0, 60), np.linspace(-3.0, 1.0, 30)])
    pubs = {"xf": [], "sb": []}
    lead = NS(px4_odom=Odometry(), xf_formation_ref=Point(0.0, 1.0, 1.0), sb_formation_ref=Point(0.0, -1.0, 1.0),
              pub_xf_formation_ref=NS(publish=lambda p: pubs["xf"].append((p.x, p.y, p.z))),
              pub_sb_formation_ref=NS(publish=lambda p: pubs["sb"].append((p.x, p.y, p.z))))
    fols = {k: NS(formation_ref=Point(1, 1, 0.5), lpf_ref_alpha=0.8, lpf_form_ref_x=None, lpf_form_ref_y=None, lpf_form_ref_z=None)
            for k in ("xf", "sb")}
    filt = {"xf": [], "sb": []}
    for xl in leader_x:
        lead.px4_odom.pose.pose.position.x = float(xl)
        ref_leader.NDPLeaderNode.pub_formation_ref_callback(lead, None)
        for k in ("xf", "sb"):
            ref_follower.FollowerNode.sub_formation_ref_callback(fols[k], Point(*pubs[k][-1]))
            f = fols[k].formation_ref
            filt[k].append((f.x, f.y, f.z))
    out["form_leader_x"] = leader_x
    out["form_xf_raw"], out["form_sb_raw"] = np.array(pubs["xf"]), np.array(pubs["sb"])
    out["form_xf_filtered"], out["form_sb_filtered"] = np.array(filt["xf"]), np.array(filt["sb"])

    # ---- 5. the leader's 1 m gate (which neighbour messages reach the MLP) ----
    ego_xy = rng.normal(size=(12, 2))
    oth_xy = ego_xy + rng.uniform(-1.2, 1.2, size=(12, 2))
    oth_xy[0] = ego_xy[0] + np.array([1.0, 0.0])  # exactly on the gate radius: "<" keeps it out
    called = []
    for i in range(12):
        m = PredXU()
        for k in range(21):
            a = Float64MultiArray(); a.data = [float(oth_xy[i, 0]), float(oth_xy[i, 1])] + [0.0] * 8
            m.x.append(a)
        od = Odometry(); od.pose.pose.position.x, od.pose.pose.position.y = float(ego_xy[i, 0]), float(ego_xy[i, 1])
        obs = NS(update=lambda other, ego: "mlp")
        ld = NS(nmpc_x_ref=np.zeros((21, 10)), px4_odom=od, downwash_observer=obs, disturb_force=None)
        ref_leader.NDPLeaderNode.sub_xf_pred_callback(ld, m)
        called.append(isinstance(ld.disturb_force, str))
    out["gate_ego_xy"], out["gate_other_xy"], out["gate_mlp_called"] = ego_xy, oth_xy, np.array(called)

    # ---- 6. the sliding 101-point list: reset + 130 control ticks on eight_low, jitter-free 50 Hz clock ----
    tc = traj_coefficients(TrajCoefficients, "eight_low")
    pub = NMPCRefPublisher()
    CLOCK[0] = 100.0
    pub.reset(tc, Time(CLOCK[0]))
    xr0, ur0 = pub.get_nmpc_ref_from_long_list()
    out["longlist_x_init"], out["longlist_u_init"] = np.array(pub.x_long_list), np.array(pub.u_long_list)  # the 101 points after reset
    seq_x, seq_u, new_x, new_u = [xr0], [ur0], [], []
    for j in range(130):
        xr_j, ur_j = pub.get_nmpc_pts(Time(CLOCK[0]))
        seq_x.append(xr_j); seq_u.append(ur_j)
        new_x.append(pub.x_long_list[-1]); new_u.append(pub.u_long_list[-1])  # the point this tick appended
        CLOCK[0] += 0.02
    out["longlist_xr"], out["longlist_ur"] = np.stack(seq_x), np.stack(seq_u)
    out["longlist_new_x"], out["longlist_new_u"] = np.stack(new_x), np.stack(new_u)
    # the hover reference before any trajectory (gen_fix_pt_ref, pt_publisher.py:40-55)
    xf, uf = NMPCRefPublisher().gen_fix_pt_ref(sim.mul_odom[0])
    out["fixpt_xr"], out["fixpt_ur"] = xf, uf

    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wire_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("xf raw values:", np.unique(out["form_xf_raw"], axis=0), "gate:", out["gate_mlp_called"])


if __name__ == "__main__":
    main()
