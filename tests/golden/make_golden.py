"""Extracts the integer tables embedded in the reference (SubrosaDG src/Solver/SimulationControl.cpp) into
tests/golden/reference_tables.json.  Run in the development container (needs /root/reference); the JSON is committed
because /root/reference does not exist on the GPU box.

These literals are the only "golden vectors" the reference holds for the hot path (it ships no tests): quadrature point
counts per order (:268-273), local face -> corner maps (:177-216), the right-side face-point permutation per rotation
(:381-523) and the parent node ids of every local face for the high-order node numbering (:525-887).
"""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/Solver/SimulationControl.cpp"
text = open(SRC).read()
lines = text.split("\n")


def function_body(name):
    """Source lines of the (first) definition of `name`, located by name and closed at the next top-level 'template <'."""
    start = next(i for i, l in enumerate(lines) if re.search(r"\b" + name + r"\(", l) and "return" not in l and ";" not in l.split("(")[0])
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("template <") or lines[i].startswith("inline constexpr std::array<int, 12>"))
    return start + 1, lines[start:end]


def scan(name):
    first, body = function_body(name)
    out = {}
    etype, parent, order, case = None, None, None, None
    pending = None
    for l in body:
        if pending is not None:  # continuation of a multi-line `return {...};`
            pending += " " + l.strip()
            if "};" not in l:
                continue
            l, pending = pending, None
        m = re.search(r"ElementType == ElementEnum::(\w+)", l)
        if m:
            etype, parent, order, case = m.group(1), None, None, None
        m = re.search(r"parent == getElementGmshTypeNumber<ElementEnum::(\w+)", l)
        if m:
            parent, order, case = m.group(1), None, None
        m = re.search(r"PolynomialOrder == (\d)", l)
        if m:
            order, case = int(m.group(1)), None
        m = re.search(r"case (\d+):", l)
        if m:
            case = int(m.group(1))
        if "return {" in l and "};" not in l:
            pending = l.strip()
            continue
        m = re.search(r"return \{(.*)\};", l)
        if m and etype is not None and m.group(1).strip():
            val = [int(x) for x in m.group(1).split(",") if x.strip()]
            key = etype + ("" if parent is None else f"/in{parent}") + ("" if order is None else f"/P{order}") + ("" if case is None else f"/case{case}")
            out[key] = val
    return {"line": first, "tables": out}


golden = {"source": "src/Solver/SimulationControl.cpp", "quadrature_number": {}}
for m in re.finditer(r"inline constexpr std::array<int, 12> k(\w+)QuadratureNumber\{([^}]*)\}", text):
    golden["quadrature_number"][m.group(1)] = [int(x) for x in m.group(2).split(",")]
golden["per_adjacency_node_index"] = scan("getElementPerAdjacencyNodeIndex")
golden["adjacency_quadrature_sequence"] = scan("getAdjacencyElementQuadratureSequence")
golden["adjacency_view_node_parent_sequence"] = scan("getAdjacencyElementViewNodeParentSequence")
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tables.json")
json.dump(golden, open(out, "w"), indent=0, sort_keys=True)
print({k: (len(v["tables"]) if isinstance(v, dict) and "tables" in v else len(v)) for k, v in golden.items() if k != "source"})
