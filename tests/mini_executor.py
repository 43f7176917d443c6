"""A minimal stand-in for the ComfyUI executor (test infrastructure): runs a prompt in ComfyUI's API format
{node_id: {"class_type": str, "inputs": {name: value | [src_id, output_index]}}} against a NODE_CLASS_MAPPINGS dict the
way the host does: validate the inputs against INPUT_TYPES(), fill widget defaults, instantiate the
class, call getattr(obj, FUNCTION)(**inputs) and expect a tuple matching RETURN_TYPES."""
from typing import Any, Dict


class PromptError(Exception):
    pass


def _validate(class_type: str, spec: Dict[str, Any], inputs: Dict[str, Any]) -> Dict[str, Any]:
    out = {}
    known = {}
    for section in ("required", "optional"):
        for name, decl in (spec.get(section) or {}).items():
            known[name] = (section, decl)
    for name in inputs:
        if name not in known:
            raise PromptError(f"{class_type}: unknown input {name!r}")
    for name, (section, decl) in known.items():
        typ = decl[0]
        opts = decl[1] if len(decl) > 1 and isinstance(decl[1], dict) else {}
        if name not in inputs:
            if section == "required" and "default" not in opts:
                raise PromptError(f"{class_type}: required input {name!r} missing")
            if "default" in opts:
                out[name] = opts["default"]
            continue
        v = inputs[name]
        if isinstance(typ, (list, tuple)):  # combo
            if v not in typ:
                raise PromptError(f"{class_type}.{name}: {v!r} not in {list(typ)}")
        elif typ == "INT":
            if not isinstance(v, int) or isinstance(v, bool):
                raise PromptError(f"{class_type}.{name}: expected INT, got {type(v).__name__}")
            if ("min" in opts and v < opts["min"]) or ("max" in opts and v > opts["max"]):
                raise PromptError(f"{class_type}.{name}: {v} outside [{opts.get('min')}, {opts.get('max')}]")
        elif typ == "FLOAT":
            if not isinstance(v, (int, float)) or isinstance(v, bool):
                raise PromptError(f"{class_type}.{name}: expected FLOAT, got {type(v).__name__}")
            if ("min" in opts and v < opts["min"]) or ("max" in opts and v > opts["max"]):
                raise PromptError(f"{class_type}.{name}: {v} outside [{opts.get('min')}, {opts.get('max')}]")
        elif typ == "BOOLEAN":
            if not isinstance(v, bool):
                raise PromptError(f"{class_type}.{name}: expected BOOLEAN, got {type(v).__name__}")
        elif typ == "STRING":
            if not isinstance(v, str):
                raise PromptError(f"{class_type}.{name}: expected STRING, got {type(v).__name__}")
        out[name] = v
    return out


def execute(prompt: Dict[str, Dict[str, Any]], mappings: Dict[str, type]) -> Dict[str, tuple]:
    """Executes every node once in dependency order; returns {node_id: output tuple}."""
    done: Dict[str, tuple] = {}
    visiting = set()

    def run(nid: str) -> tuple:
        if nid in done:
            return done[nid]
        if nid in visiting:
            raise PromptError(f"cycle through node {nid}")
        visiting.add(nid)
        node = prompt[nid]
        ct = node["class_type"]
        if ct not in mappings:
            raise PromptError(f"node type {ct!r} is not registered")
        cls = mappings[ct]
        spec = cls.INPUT_TYPES()
        raw, links = {}, {}
        for name, v in node.get("inputs", {}).items():
            if isinstance(v, list) and len(v) == 2 and isinstance(v[0], str) and v[0] in prompt:
                links[name] = v
            else:
                raw[name] = v
        known = {**(spec.get("required") or {}), **(spec.get("optional") or {})}
        for name, (src, idx) in links.items():
            if name not in known:
                raise PromptError(f"{ct}: unknown input {name!r}")
            src_cls = mappings[prompt[src]["class_type"]]
            want = known[name][0]
            if src_cls.RETURN_TYPES[idx] != want:
                raise PromptError(f"{ct}.{name}: link carries {src_cls.RETURN_TYPES[idx]}, socket wants {want}")
        kwargs = _validate(ct, {k: {n: d for n, d in (spec.get(k) or {}).items() if n not in links} for k in ("required", "optional")}, raw)
        for name, (src, idx) in links.items():
            kwargs[name] = run(src)[idx]
        res = getattr(cls(), cls.FUNCTION)(**kwargs)
        if not isinstance(res, tuple) or len(res) != len(cls.RETURN_TYPES):
            raise PromptError(f"{ct}: returned {type(res).__name__} instead of a {len(cls.RETURN_TYPES)}-tuple")
        visiting.discard(nid)
        done[nid] = res
        return res

    for nid in prompt:
        run(nid)
    return done
