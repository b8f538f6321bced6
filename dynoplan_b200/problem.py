"""Planning problem description (environment + start/goal).

Mirror of ``dynobench::Problem::read_from_yaml`` (dynobench/src/motions.cpp:227-283)
and of the obstacle conventions of ``load_env`` (dynobench/src/motions.cpp:656-688):
2-D boxes get z-size 1 centred at z = 0, spheres keep ``size[0]`` as radius.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import yaml

BOX, SPHERE = 0, 1


@dataclass
class Obstacle:
    type: int  # BOX | SPHERE
    center: List[float]  # always 3 entries
    size: List[float]  # always 3 entries (box: full sizes, sphere: [r, 0, 0])


@dataclass
class Problem:
    name: str = ""
    robotType: str = ""
    start: List[float] = field(default_factory=list)
    goal: List[float] = field(default_factory=list)
    p_lb: List[float] = field(default_factory=list)
    p_ub: List[float] = field(default_factory=list)
    obstacles: List[Obstacle] = field(default_factory=list)

    @staticmethod
    def from_dict(env: dict) -> "Problem":
        p = Problem()
        p.name = env.get("name", "")
        robot = env["robots"][0]
        p.robotType = robot["type"]
        p.start = [float(v) for v in robot["start"]]
        p.goal = [float(v) for v in robot["goal"]]
        p.p_lb = [float(v) for v in env["environment"]["min"]]
        p.p_ub = [float(v) for v in env["environment"]["max"]]
        if len(p.p_lb) != len(p.p_ub) or len(p.p_lb) > 3:
            raise ValueError("environment.min/max must have equal size <= 3")
        for obs in env["environment"].get("obstacles") or []:
            size = [float(v) for v in obs["size"]]
            center = [float(v) for v in obs["center"]]
            t = obs["type"]
            if t == "box":
                # load_env: z-size 1 / z-centre 0 when the box is given in 2-D
                sz = [size[0], size[1], size[2] if len(size) == 3 else 1.0]
                c = [center[0], center[1], center[2] if len(size) == 3 else 0.0]
                p.obstacles.append(Obstacle(BOX, c, sz))
            elif t == "sphere":
                c = [center[0], center[1], center[2] if len(center) == 3 else 0.0]
                p.obstacles.append(Obstacle(SPHERE, c, [size[0], 0.0, 0.0]))
            else:
                raise RuntimeError("Unknown obstacle type! --" + str(t))  # motions.cpp:682
        # quaternions of quad3d start/goal are normalised (motions.cpp:277-281)
        if p.robotType.startswith("quad3d") and not p.robotType.startswith("quad3dpayload"):
            for v in (p.start, p.goal):
                n = math.sqrt(sum(a * a for a in v[3:7]))
                v[3:7] = [a / n for a in v[3:7]]
        return p

    @staticmethod
    def from_yaml(path: str) -> "Problem":
        with open(path) as f:
            return Problem.from_dict(yaml.safe_load(f))
