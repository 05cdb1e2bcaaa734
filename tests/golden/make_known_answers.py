"""Derivation of tests/golden/known_answers.json (closed forms, no reference code involved).

extrapolation row 0: N_g(node 0 scaled by sqrt(3)) = (1 -/+ s*sqrt(3))-products / 8,
  a = (1+sqrt3)/2, b = (1-sqrt3)/2  ->  a^3, a^2 b, a b^2, a^2 b, a^2 b, a b^2, b^3, a b^2.
unit cube K00 (ux at node 1 with itself, full integration, edge 1):
  (lambda + 2G)/9 + G/9 + G/9,  lambda = E nu/((1+nu)(1-2nu)), G = E/(2(1+nu)).
bfs_2x2x2: start = node 0 (first node with one incident element); its neighbours in the
  CHEXA order of element 0 are 1, 4, 3, 9, 10, 13, 12 (node index = i + 3(j + 3k)).
"""
import json
import math

a, b = (1 + math.sqrt(3)) / 2, (1 - math.sqrt(3)) / 2
row0 = [a**3, a * a * b, a * b * b, a * a * b, a * a * b, a * b * b, b**3, a * b * b]
E, nu = 210000.0, 0.3
lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
print(json.dumps({"row0": row0, "K00": (lam + 2 * G) / 9 + 2 * G / 9, "lambda": lam, "G": G}, indent=1))
