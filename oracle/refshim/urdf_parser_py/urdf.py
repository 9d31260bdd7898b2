"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `urdf_parser_py` package.

The reference (`/root/reference/differentiable_robot_model/urdf_utils.py:9,14`) imports
`urdf_parser_py.urdf.URDF` purely to *parse XML*; the package is not installed in this image and
there is no network.  This module exposes exactly the attribute surface the reference reads
(`urdf_utils.py:17-26,43-75,85-108`) so that `/root/reference` can be imported, unmodified, in
this container to (a) pin `oracle/` and (b) generate the golden vectors under `tests/golden/`
(`tests/golden/make_golden.py`).  It performs no arithmetic and is never imported by the product.

Behaviour mirrored from urdf_parser_py (PyPI `urdf-parser-py`, unpinned in the reference's
requirements.txt:5):
  * links and joints are kept in document order;
  * `<origin>` attributes `xyz` / `rpy` default to [0, 0, 0] when missing;
  * `link.inertial` is None when the element is absent;
  * `joint.dynamics` / `joint.limit` / `joint.axis` are None when the element is absent
    (the reference turns `None.damping` into damping 0 via `except AttributeError`).
"""
import xml.etree.ElementTree as ET


def _vec(text, n=3):
    if text is None:
        return [0.0] * n
    vals = [float(v) for v in text.split()]
    assert len(vals) == n, f"expected {n} numbers, got {text!r}"
    return vals


class Pose:
    def __init__(self, elem=None):
        self.xyz = _vec(elem.get("xyz") if elem is not None else None)
        self.rpy = _vec(elem.get("rpy") if elem is not None else None)

    @property
    def position(self):
        return self.xyz

    @property
    def rotation(self):
        return self.rpy


class Inertia:
    def __init__(self, elem):
        for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"):
            setattr(self, k, float(elem.get(k, 0.0)))


class Inertial:
    def __init__(self, elem):
        self.origin = Pose(elem.find("origin"))
        self.mass = float(elem.find("mass").get("value"))
        self.inertia = Inertia(elem.find("inertia"))


class Link:
    def __init__(self, elem):
        self.name = elem.get("name")
        inertial = elem.find("inertial")
        self.inertial = Inertial(inertial) if inertial is not None else None


class JointLimit:
    def __init__(self, elem):
        self.effort = float(elem.get("effort", 0.0))
        self.velocity = float(elem.get("velocity", 0.0))
        self.lower = float(elem.get("lower", 0.0))
        self.upper = float(elem.get("upper", 0.0))


class JointDynamics:
    def __init__(self, elem):
        self.damping = float(elem.get("damping", 0.0))
        self.friction = float(elem.get("friction", 0.0))


class Joint:
    def __init__(self, elem):
        self.name = elem.get("name")
        self.type = elem.get("type")
        self.parent = elem.find("parent").get("link")
        self.child = elem.find("child").get("link")
        self.origin = Pose(elem.find("origin"))
        axis = elem.find("axis")
        self.axis = _vec(axis.get("xyz")) if axis is not None else None
        limit = elem.find("limit")
        self.limit = JointLimit(limit) if limit is not None else None
        dyn = elem.find("dynamics")
        self.dynamics = JointDynamics(dyn) if dyn is not None else None


class URDF:
    def __init__(self):
        self.name = ""
        self.links = []
        self.joints = []

    @classmethod
    def from_xml_file(cls, path):
        root = ET.parse(path).getroot()
        robot = cls()
        robot.name = root.get("name", "")
        robot.links = [Link(e) for e in root.findall("link")]
        robot.joints = [Joint(e) for e in root.findall("joint")]
        return robot
