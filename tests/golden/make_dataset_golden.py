#!/usr/bin/env python
"""Writes the fixtures of the offline dataset pipeline from the reference tree (run in the build container):

  tests/golden/dataset/g1_29dof_kinematic.urdf     the reference URDF (g1_model/urdf/g1_29dof_rev_1_0.urdf) with everything but the
                                                    kinematic tree removed (links, joints: parent / child / origin / axis / limit)
  tests/golden/dataset/walk1_rows_110_265.csv       lines 110..264 of datasets/walk1_subject1.csv, verbatim -- the slice the shipped
                                                    motions/custom_motion.npz was converted from
  tests/golden/dataset/data_convert_output.npz      output of the UNMODIFIED motions/data_convert.py run HERE on those inputs, with
                                                    oracle.dataset_oracle's forward kinematics standing in for the absent Pinocchio

The shipped motions/custom_motion.npz (the reference's own output, made with the real Pinocchio) is already a fixture:
tests/golden/clips_full/custom_motion.npz.
"""
import os
import runpy
import sys
import tempfile
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("AMP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "dataset")


def strip_urdf(src, dst):
    robot = ET.parse(src).getroot()
    slim = ET.Element("robot", {"name": robot.get("name", "robot")})
    for link in robot.findall("link"):
        ET.SubElement(slim, "link", {"name": link.get("name")})
    for j in robot.findall("joint"):
        e = ET.SubElement(slim, "joint", {"name": j.get("name"), "type": j.get("type")})
        for tag in ("origin", "parent", "child", "axis", "limit"):
            c = j.find(tag)
            if c is not None:
                ET.SubElement(e, tag, dict(c.attrib))
    ET.indent(slim)
    ET.ElementTree(slim).write(dst, encoding="utf-8", xml_declaration=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    urdf = os.path.join(OUT, "g1_29dof_kinematic.urdf")
    strip_urdf(os.path.join(REF, "g1_model", "urdf", "g1_29dof_rev_1_0.urdf"), urdf)
    csv = os.path.join(OUT, "walk1_rows_110_265.csv")
    with open(os.path.join(REF, "datasets", "walk1_subject1.csv")) as f:
        lines = f.readlines()[110:265]
    with open(csv, "w") as f:
        f.writelines(lines)
    # the unmodified reference tool, Pinocchio replaced by the restated FK
    from oracle import dataset_oracle

    sys.modules["pinocchio"] = dataset_oracle.make_pinocchio_stub()
    out = os.path.join(OUT, "data_convert_output.npz")
    with tempfile.TemporaryDirectory() as tmp:
        argv, sys.argv = sys.argv, ["data_convert.py", "--csv", csv, "--urdf", urdf, "--meshes", tmp, "--output", out]
        try:
            runpy.run_path(os.path.join(REF, "motions", "data_convert.py"), run_name="__main__")
        finally:
            sys.argv = argv
            del sys.modules["pinocchio"]
    d = np.load(out)
    print({k: (d[k].shape, str(d[k].dtype)) for k in d.files})


if __name__ == "__main__":
    main()
