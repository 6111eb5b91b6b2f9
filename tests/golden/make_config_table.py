"""Writes tests/golden/reference_configs.json: the numbers of the reference's shipped config files that
ganecdotes_b200/configs.py restates (configs/segmentors/hfc_with_swav*_config.py, configs/models/*.py and the
model-name table of configs/mapper.py).  Run in the build container (reads /root/reference):

    python tests/golden/make_config_table.py
"""
import ast
import json
import os
import re

REF = os.environ.get("GX_REFERENCE", "/root/reference")


def seg_config(path):
    """the segmentor config files are plain Python (no imports): execute them"""
    ns = {}
    exec(compile(open(path).read(), path, "exec"), ns)
    a = ns["hfc_prep_args"]
    sw = {k: v for k, v in a["swav_args"].items()}
    return dict(perturb_args=a["perturb_args"], swav_args=sw, sinkhorn_args=a["sinkhorn_args"],
                layer_hf_dim=a["layer_hf_dim"], seg_args=ns["seg_args"], n_hfc_layers=ns["n_hfc_layers"])


def model_config(path):
    """the model config files import the reference's lib: read the literals with ast instead"""
    out = {}
    for node in ast.parse(open(path).read()).body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name in ("num_latents_for_mean", "truncation", "image_size", "latent_dim", "is_baggan"):
                out[name] = ast.literal_eval(node.value)
            if name == "gen_args":
                kw = {}
                for k in node.value.keywords:
                    try:
                        kw[k.arg] = ast.literal_eval(k.value)
                    except Exception:
                        kw[k.arg] = out.get(ast.unparse(k.value), ast.unparse(k.value))   # e.g. size=image_size
                out[name] = kw
    return out


def main():
    mapper = open(os.path.join(REF, "configs", "mapper.py")).read()
    models = dict(re.findall(r"'([\w-]+)':\s+os\.path\.join\(CONFIGS_DIR, 'models', '([\w.]+)'\)", mapper))
    segs = dict(re.findall(r"'([\w-]+)':\s+os\.path\.join\(CONFIGS_DIR, 'segmentors', '([\w.]+)'\)", mapper))
    table = {"models": {}, "segmentors": {}}
    for name, f in sorted(models.items()):
        p = os.path.join(REF, "configs", "models", f)
        if os.path.exists(p):
            table["models"][name] = model_config(p)
    for name, f in sorted(segs.items()):
        if name.startswith("hfc_with_swav"):
            table["segmentors"][name] = seg_config(os.path.join(REF, "configs", "segmentors", f))
    table["baselines"] = {}
    for name in ("hfc_with_simclr", "hfc_kmeans"):          # plain Python too
        ns = {}
        path = os.path.join(REF, "configs", "segmentors", segs[name])
        exec(compile(open(path).read(), path, "exec"), ns)
        a = dict(ns["hfc_prep_args"])
        a.pop("train", None)
        table["baselines"][name] = dict(hfc_prep_args=a, seg_args=ns["seg_args"], n_hfc_layers=ns["n_hfc_layers"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_configs.json")
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
    print(out, len(table["models"]), "models", len(table["segmentors"]), "segmentor configs")


if __name__ == "__main__":
    main()
