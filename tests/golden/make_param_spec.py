"""Writes oracle/param_spec_{camliraft,camlipwc,camliraft_l}.json: parameter/buffer names -> shapes of the
REFERENCE models (built from /root/reference through tests/golden/ref_harness.py; build
container only).  The oracle and the product both key their seeded weights on these names,
so a test that the product's state_dict matches the spec is a test of checkpoint
compatibility with the reference (checkpoint row).

    python tests/golden/make_param_spec.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def main():
    models = rh.load_reference()
    nets = {
        "camliraft": models.camliraft.CamLiRAFT(rh.camliraft_cfg()),
        "camlipwc": models.camlipwc.CamLiPWC(rh.camlipwc_cfg()),
        "camliraft_l": models.camliraft_l.CamLiRAFT_L(rh.camliraft_l_cfg()),
    }
    for name, net in nets.items():
        spec = {k: list(v.shape) for k, v in net.state_dict().items()}
        path = os.path.join(ROOT, "oracle", "param_spec_%s.json" % name)
        with open(path, "w") as f:
            json.dump(spec, f, indent=0, sort_keys=True)
        print(path, len(spec), sum(int(__import__("numpy").prod(s)) for s in spec.values()))


if __name__ == "__main__":
    main()
