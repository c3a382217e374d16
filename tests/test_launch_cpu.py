"""ilswiss_b200.launch (the run_experiment.py workflow without editing or copying the reference): variant expansion through
the reference's own build_nested_variant_generator, log-directory redirection, one child per variant running the named
experiment script with runpy.  The real scripts need gym / envpool / MuJoCo (absent here), so the child
script of this test is a probe that records what a real script would see."""
import json
import os
import sys

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not importable here")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = '''
import argparse, json, os, sys, yaml
ap = argparse.ArgumentParser(); ap.add_argument("-e"); ap.add_argument("-g", type=int)
a = ap.parse_args()
from rlkit.launchers import config
v = yaml.safe_load(open(a.e))
import rlkit.torch.algorithms.sac.sac_alpha as m
out = dict(seed=v["seed"], exp_id=v["exp_id"], gpu=a.g, log_dir=config.LOCAL_LOG_DIR, trainer_module=m.SoftActorCritic.__module__,
           net_size=v["net_size"])
open(os.path.join(config.LOCAL_LOG_DIR, "probe_%d.json" % v["exp_id"]), "w").write(json.dumps(out))
'''


def test_launcher_expands_variants_and_runs_the_script_per_variant(tmp_path):
    import yaml

    script = tmp_path / "probe_exp_script.py"
    script.write_text(PROBE)
    spec = dict(meta_data=dict(script_path=str(script), exp_name="probe", description="", num_workers=2, using_gpus=True),
                variables=dict(seed=[0, 1, 2]), constants=dict(net_size=256, ilswiss_b200=False))
    spec_path = tmp_path / "spec.yaml"
    spec_path.write_text(yaml.dump(spec))
    log_dir = tmp_path / "logs"
    # this image lacks matplotlib / seaborn / gtimer / gym, which rlkit imports at module level: the child interpreters get
    # the test shim through a sitecustomize module (test infrastructure; a real deployment has those packages)
    (tmp_path / "sitecustomize.py").write_text("from oracle import ref_shim\nref_shim.install()\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT, os.environ.get("PYTHONPATH", "")]))
    import subprocess

    rc = subprocess.call([sys.executable, "-m", "ilswiss_b200.launch", "-e", str(spec_path), "-g", "3", "--reference",
                          ref_shim.REFERENCE_ROOT, "--log-dir", str(log_dir)], env=env, cwd=str(tmp_path))
    assert rc == 0
    got = sorted((json.loads(open(os.path.join(log_dir, f)).read()) for f in os.listdir(log_dir) if f.startswith("probe_")),
                 key=lambda d: d["exp_id"])
    assert [g["seed"] for g in got] == [0, 1, 2] and all(g["gpu"] == 3 and g["net_size"] == 256 for g in got)
    assert all(os.path.samefile(g["log_dir"], log_dir) for g in got)                 # config.LOCAL_LOG_DIR redirected
    assert all(g["trainer_module"].startswith("rlkit.") for g in got)               # ilswiss_b200: false -> pure reference
    vdirs = [d for d in os.listdir(log_dir) if d.startswith("variants-for-probe")]
    assert len(vdirs) == 1
