"""Signatures of the reference functions on the drop-in boundary -> tests/golden/reference_signatures.json.

Test infrastructure, run in the authoring container only (it reads /root/reference).  The reference package does not
import as a whole, so the signatures are taken from the source with `ast`: parameter names in order, defaults as
Python literals (module-level constants such as MASK_Z or nb_nodes are resolved to their literal values)."""
import ast
import json
import os
import sys

REF = os.environ.get("DISCO_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_signatures.json")

FUNCTIONS = {
    "disco_theque/speech_enhancement/tango.py": ["offline_tango", "get_mask", "concatenate_signals", "get_z_for_mask", "reshape_mask"],
    "disco_theque/speech_enhancement/get_z_signals.py": ["offline_tango"],
    "disco_theque/se_utils/internal_formulas.py": ["get_filter_type", "intern_filter", "spatial_correlation_matrix"],
    "disco_theque/dnn/utils.py": ["tf_mask"],
    "disco_theque/sigproc_utils.py": ["tf_mask", "vad_oracle_batch"],
    "disco_theque/math_utils.py": ["my_stft", "my_istft"],
    "disco_theque/metrics.py": ["snr", "delta_snr", "sd", "fw_snr", "fw_sd", "si_sdr"],
    "disco_theque/speech_enhancement/utils.py": ["prepare_data"],
}


def literal(node, consts):
    if isinstance(node, ast.Name) and node.id in consts:
        return consts[node.id]
    if (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "len" and len(node.args) == 1
            and isinstance(node.args[0], ast.Name) and node.args[0].id in consts):
        return len(consts[node.args[0].id])            # nb_nodes = len(nb_ch), tango.py:30-31
    if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "array"
            and len(node.args) == 1):
        return ast.literal_eval(node.args[0])          # nb_ch = np.array([4, 4, 4, 4])
    return ast.literal_eval(node)


def main():
    out = {}
    for rel, names in FUNCTIONS.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        consts = {}
        for node in tree.body:                      # module-level NAME = literal
            if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                try:
                    consts[node.targets[0].id] = literal(node.value, consts)
                except Exception:
                    pass
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                a = node.args
                pos = [x.arg for x in a.posonlyargs + a.args]
                defaults = [None] * (len(pos) - len(a.defaults)) + [literal(d, consts) for d in a.defaults]
                has_default = [False] * (len(pos) - len(a.defaults)) + [True] * len(a.defaults)
                out["%s:%s" % (rel, node.name)] = {
                    "line": node.lineno,
                    "params": [{"name": p, "has_default": h, "default": d} for p, h, d in zip(pos, has_default, defaults)],
                    "varargs": a.vararg.arg if a.vararg else None, "varkw": a.kwarg.arg if a.kwarg else None}
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "signatures")


if __name__ == "__main__":
    sys.exit(main())
