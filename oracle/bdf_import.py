"""TEST INFRASTRUCTURE, NOT PRODUCT CODE: a second, independent restatement of the reference's Nastran
importer, used only to cross-check stan_b200/host/bdf.cpp on generated decks (tests/test_host_codec.py).

Follows Database.ReadNastranMesh (/root/reference/src/STAN_Database/Database.cs:39-111), the
fixed-width GRID parser Node(string) (Node.cs:25-80) and the whitespace-split CHEXA parser
Element(string) (Element.cs:35-73), including what is unusual about them: only `-` exponents are
patched (`+` is replaced into a discarded string), blank fields are dropped rather than kept as
columns, elements keep however many integers followed, duplicate IDs are import errors, and any
line that merely contains "CHEXA" starts an element.  PARITY UNPINNED like the rest of oracle/.
"""
from __future__ import annotations

import re

_INT = re.compile(r"^[ \t\n\v\f\r]*[+-]?[0-9]+[ \t\n\v\f\r]*$")
_DBL = re.compile(r"^[ \t\n\v\f\r]*[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?[ \t\n\v\f\r]*$")


def _int_parse(s: str) -> int:                       # int.Parse (NumberStyles.Integer)
    if not _INT.match(s):
        raise ValueError(s)
    v = int(s)
    if not -2**31 <= v <= 2**31 - 1:
        raise OverflowError(s)
    return v


def _double_parse(s: str) -> float:                  # double.Parse(s, InvariantCulture), deck-sized subset
    if not _DBL.match(s):
        raise ValueError(s)
    return float(s)


def parse_grid(line: str):
    data = []
    for i in range(len(line) // 8):
        text = line[8 * i:8 * i + 8].replace(" ", "")
        if text.strip(" \t\n\v\f\r") == "":
            continue
        if "e" not in text and "E" not in text:
            if "-" in text[1:]:
                text = "-" + text[1:].replace("-", "e-") if text[0] == "-" else text.replace("-", "e-")
        if text[0] == ".":
            text = "0" + text
        data.append(text)
    return _int_parse(data[1]), _double_parse(data[2]), _double_parse(data[3]), _double_parse(data[4])


def parse_element(text: str):
    data = re.split(r"[ \t\n\v\f\r]+", text)
    eid, pid = _int_parse(data[1]), _int_parse(data[2])
    nlist = []
    for tok in data[3:]:
        tok = tok.replace("+", "")
        try:
            nlist.append(_int_parse(tok))
        except (ValueError, OverflowError):
            pass
    etype = {"CHEXA": "HEX8_G2", "CPENTA": "PENTA6_G2", "CTETRA": "TET4_G2"}.get(data[0], "")
    return eid, pid, nlist, etype


def read_nastran_mesh(text: str):
    """Returns (nodes {id: (x, y, z)} in insertion order, elements {id: (pid, nlist, type)}, n_errors)."""
    lines = re.split(r"\r\n|\n|\r", text)
    if lines and lines[-1] == "":
        lines.pop()                                  # File.ReadAllLines has no entry after a final newline
    nodes, elems, errors = {}, {}, 0
    i = 0
    while i < len(lines):
        if not lines[i].startswith("$"):
            if "CHEXA" in lines[i]:
                temp = lines[i]
                j = i + 1
                while j < len(lines) and (lines[j].startswith("+") or lines[j].startswith(" ")):
                    temp += lines[j]
                    i = j
                    j += 1
                try:
                    eid, pid, nlist, etype = parse_element(temp)
                    if eid in elems:
                        raise KeyError(eid)
                    elems[eid] = (pid, nlist, etype)
                except (ValueError, OverflowError, IndexError, KeyError):
                    errors += 1
            if lines[i].startswith("GRID"):
                try:
                    nid, x, y, z = parse_grid(lines[i])
                    if nid in nodes:
                        raise KeyError(nid)
                    nodes[nid] = (x, y, z)
                except (ValueError, OverflowError, IndexError, KeyError):
                    errors += 1
        i += 1
    return nodes, elems, errors


def parse_bc_text(text: str):
    """BOX_BC.Paste_Click (/root/reference/src/STAN_PrePost/BOX_BC.xaml.cs:228-270): rows "NID x y z" split by
    ',' then ' ' then TAB (first split that yields exactly four fields), unparsable rows skipped, nothing read
    from a text of fewer than two lines."""
    rows = []
    lines = text.split("\n")
    if len(lines) <= 1:
        return rows
    for s in lines:
        f = s.split(",")
        if len(f) != 4:
            f = s.split(" ")
            if len(f) != 4:
                f = s.split("\t")
        if len(f) != 4:
            continue
        try:
            rows.append((_int_parse(f[0]), _double_parse(f[1]), _double_parse(f[2]), _double_parse(f[3])))
        except (ValueError, OverflowError):
            pass
    return rows


def apply_bc(rows, known_ids):
    """BoundaryCondition.Add (BoundaryCondition.cs:87-98): unknown nodes dropped, a repeated node raises."""
    out, seen = [], set()
    for nid, x, y, z in rows:
        if nid not in known_ids:
            continue
        if nid in seen:
            raise KeyError(nid)
        seen.add(nid)
        out.append((nid, [x, y, z]))
    return out
