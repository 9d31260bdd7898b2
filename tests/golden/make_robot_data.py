#!/usr/bin/env python
"""Derive the kinematic/inertial-only robot descriptions shipped in
`differentiable_robot_model_b200/robot_data/` from the URDFs of the reference's `diff_robot_data/`.

Run in the build container only (needs /root/reference):

    python tests/golden/make_robot_data.py

For every well-formed URDF under /root/reference/diff_robot_data this writes a *reduced* URDF that
keeps, in document order, only what the reference's loader reads
(`differentiable_robot_model/urdf_utils.py:28-126`): `<link name>` + `<inertial>` (origin, mass,
inertia) and `<joint name type>` + parent / child / origin / axis / limit / dynamics / mimic.
Visual, collision, material, gazebo, transmission and mesh content is dropped (the library never
reads it).  The output is a canonical serialisation (one line per link / joint, fixed child and attribute order);
attribute VALUES are copied verbatim so every parsed float is bit-identical.
`fetch.urdf` is not well-formed XML in the reference (unbound `sensor:` prefix) and is skipped.
"""
import os
import sys
import xml.etree.ElementTree as ET

SRC = "/root/reference/diff_robot_data"
DST = os.path.join(os.path.dirname(__file__), "..", "..", "differentiable_robot_model_b200", "robot_data")

LINK_KEEP = {"inertial": ("mass", "origin", "inertia")}
JOINT_KEEP = ("origin", "axis", "parent", "child", "limit", "dynamics", "mimic")


def _attrs(e):
    """Attributes in a canonical (reverse-alphabetical) order; values verbatim."""
    return "".join(f' {k}="{e.attrib[k]}"' for k in sorted(e.attrib, reverse=True))


def reduce_urdf(src_path):
    root = ET.parse(src_path).getroot()
    out = ['<?xml version="1.0"?>',
           "<!-- kinematic/inertial-only description derived by tests/golden/make_robot_data.py -->",
           f'<robot name="{root.get("name", "")}">']
    for e in root:
        if e.tag == "link":
            # canonical serialisation: one line per link / joint, fixed child and attribute order
            inertial = e.find("inertial")
            if inertial is None:
                out.append(f'<link name="{e.get("name")}"/>')
            else:
                kept = "".join(f"<{tag}{_attrs(inertial.find(tag))}/>" for tag in LINK_KEEP["inertial"]
                               if inertial.find(tag) is not None)
                out.append(f'<link name="{e.get("name")}"><inertial>{kept}</inertial></link>')
        elif e.tag == "joint":
            kept = "".join(f"<{tag}{_attrs(e.find(tag))}/>" for tag in JOINT_KEEP if e.find(tag) is not None)
            out.append(f'<joint type="{e.get("type")}" name="{e.get("name")}">{kept}</joint>')
    out.append("</robot>")
    return "\n".join(out) + "\n"


def main():
    n = 0
    for dirpath, _, files in os.walk(SRC):
        for f in sorted(files):
            if not f.endswith(".urdf"):
                continue
            src = os.path.join(dirpath, f)
            rel = os.path.relpath(src, SRC)
            try:
                text = reduce_urdf(src)
            except ET.ParseError as exc:
                print(f"skip {rel}: {exc}", file=sys.stderr)
                continue
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "w") as fh:
                fh.write(text)
            n += 1
            print(f"wrote {os.path.relpath(dst)} ({len(text.splitlines())} lines)")
    print(f"{n} descriptions")


if __name__ == "__main__":
    main()
