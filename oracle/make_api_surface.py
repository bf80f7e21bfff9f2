"""TEST INFRASTRUCTURE ONLY — records the public surface of the reference's training scripts (module-level functions,
classes, methods and their positional argument names, read with `ast`: nothing is imported or executed) into
tests/golden/api_surface.json, so that tests/test_api_surface.py can hold the drop-in modules to it on machines where
/root/reference does not exist.  Run in the build container:  python oracle/make_api_surface.py"""
import ast
import json
import os

REF_SRC = os.environ.get("GEMMGAN_REFERENCE_SRC", "/root/reference/src")
SCRIPTS = ["conditional_gan_cross_attention_with_film", "conditional_gan_cross_attention", "conditional_gan_film",
           "conditional_gan_img_transformer", "conditional_gan_attention", "conditional_gan_concat",
           "vanilla_gan_unconditional", "benchmark_generative_model", "multi_patch_gan_dataloader",
           "multi_patch_multi_token_gan_dataloader", "data_loader", "benchmark_gan_dataloader"]


def surface(path):
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            out[node.name] = [a.arg for a in node.args.args]
        elif isinstance(node, ast.ClassDef):
            for m in node.body:
                if isinstance(m, ast.FunctionDef):
                    out[f"{node.name}.{m.name}"] = [a.arg for a in m.args.args]
    return out


if __name__ == "__main__":
    data = {s: surface(os.path.join(REF_SRC, s + ".py")) for s in SCRIPTS}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "api_surface.json")
    with open(dst, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(dst, sum(len(v) for v in data.values()), "entries")
