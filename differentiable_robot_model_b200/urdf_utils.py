"""
URDF loading (host side only)
====================================
Own XML reader for the kinematic / inertial subset of URDF that the engine needs; mirrors the
interface of the reference's ``URDFRobotModel`` (``differentiable_robot_model/urdf_utils.py:12-126``),
which delegates the parsing to the third-party ``urdf_parser_py``.  Same per-link dictionary keys,
same defaults:

* body 0 is the root: zero origin, ``joint_type="fixed"``, zero axis (``urdf_utils.py:33-40``);
* a link without ``<inertial>`` gets mass 1, com 0, inertia eye(3) and a warning (``urdf_utils.py:114-124``);
* a movable joint without ``<dynamics>`` gets damping 0 (``urdf_utils.py:65-72``);
* ``<origin>`` attributes default to zeros; inertial ``rpy`` is ignored (``urdf_utils.py:89-97``);
* every non-"fixed" joint type (revolute, continuous, prismatic) is treated as revolute, as the
  reference does (``robot_model.py:123`` only tests ``!= "fixed"``).

All tensors are float32 (the reference is fp32-only).
"""
import xml.etree.ElementTree as ET

import torch


def _floats(text, n, default=0.0):
    if text is None:
        return [default] * n
    vals = [float(tok) for tok in text.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers in URDF attribute, got {text!r}")
    return vals


class _Joint:
    __slots__ = ("name", "type", "parent", "child", "xyz", "rpy", "axis", "limit", "damping")

    def __init__(self, elem):
        self.name = elem.get("name")
        self.type = elem.get("type")
        self.parent = elem.find("parent").get("link")
        self.child = elem.find("child").get("link")
        origin = elem.find("origin")
        self.xyz = _floats(origin.get("xyz") if origin is not None else None, 3)
        self.rpy = _floats(origin.get("rpy") if origin is not None else None, 3)
        axis = elem.find("axis")
        self.axis = _floats(axis.get("xyz"), 3) if axis is not None else None
        limit = elem.find("limit")
        self.limit = None
        if limit is not None:
            self.limit = {
                "effort": float(limit.get("effort", 0.0)),
                "lower": float(limit.get("lower", 0.0)),
                "upper": float(limit.get("upper", 0.0)),
                "velocity": float(limit.get("velocity", 0.0)),
            }
        dyn = elem.find("dynamics")
        self.damping = float(dyn.get("damping", 0.0)) if dyn is not None else None


class _Link:
    __slots__ = ("name", "mass", "com", "inertia")

    def __init__(self, elem):
        self.name = elem.get("name")
        inertial = elem.find("inertial")
        if inertial is None:
            self.mass = self.com = self.inertia = None
            return
        origin = inertial.find("origin")
        self.com = _floats(origin.get("xyz") if origin is not None else None, 3)
        self.mass = float(inertial.find("mass").get("value"))
        i = inertial.find("inertia")
        g = lambda k: float(i.get(k, 0.0))  # noqa: E731
        self.inertia = [
            [g("ixx"), g("ixy"), g("ixz")],
            [g("ixy"), g("iyy"), g("iyz")],
            [g("ixz"), g("iyz"), g("izz")],
        ]


class _Robot:
    def __init__(self, path):
        root = ET.parse(path).getroot()
        self.name = root.get("name", "")
        self.links = [_Link(e) for e in root.findall("link")]      # document order
        self.joints = [_Joint(e) for e in root.findall("joint")]   # document order


class URDFRobotModel(object):
    def __init__(self, urdf_path, device="cpu"):
        self.robot = _Robot(urdf_path)
        self._device = torch.device(device)
        self._joint_of_child = {}
        for j, joint in enumerate(self.robot.joints):
            self._joint_of_child.setdefault(joint.child, j)   # first match, like the reference's scan

    def find_joint_of_body(self, body_name):
        return self._joint_of_child.get(body_name, -1)

    def get_name_of_parent_body(self, link_name):
        jid = self.find_joint_of_body(link_name)
        return self.robot.joints[jid].parent

    def get_body_parameters_from_urdf(self, i, link):
        dev = self._device
        f32 = dict(dtype=torch.float32, device=dev)
        body = {"joint_id": i, "link_name": link.name}

        if i == 0:
            body.update(
                rot_angles=torch.zeros(3, **f32), trans=torch.zeros(3, **f32), joint_name="base_joint",
                joint_type="fixed", joint_limits=None, joint_damping=None, joint_axis=torch.zeros((1, 3), **f32),
            )
        else:
            jid = self.find_joint_of_body(link.name)
            if jid < 0:
                raise ValueError(f"link {link.name!r} is not the child of any joint (only link 0 may be the root)")
            joint = self.robot.joints[jid]
            limits, damping, axis = None, torch.zeros(1, **f32), torch.zeros((1, 3), **f32)
            if joint.type != "fixed":
                if joint.limit is None:
                    raise ValueError(f"movable joint {joint.name!r} has no <limit>")
                limits = dict(joint.limit)
                if joint.damping is not None:
                    damping = torch.tensor([joint.damping], **f32)
                if joint.axis is None:
                    raise ValueError(f"movable joint {joint.name!r} has no <axis>")
                axis = torch.tensor(joint.axis, **f32).reshape(1, 3)
            body.update(
                rot_angles=torch.tensor(joint.rpy, **f32), trans=torch.tensor(joint.xyz, **f32),
                joint_name=joint.name, joint_type=joint.type, joint_limits=limits, joint_damping=damping,
                joint_axis=axis,
            )

        if link.mass is not None:
            body["mass"] = torch.tensor([link.mass], **f32)
            body["com"] = torch.tensor(link.com, **f32).reshape(1, 3)
            body["inertia_mat"] = torch.tensor(link.inertia, **f32).unsqueeze(0)
        else:
            body["mass"] = torch.ones((1,), **f32)
            body["com"] = torch.zeros((1, 3), **f32)
            body["inertia_mat"] = torch.eye(3, **f32).unsqueeze(0)
            print(
                "Warning: No dynamics information for link: {}, setting all inertial properties to 1.".format(
                    link.name
                )
            )
        return body
