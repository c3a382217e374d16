"""Launcher: the reference's `run_experiment.py -e <exp_spec.yaml> -g <gpu>` workflow with the B200 path switched in.

    python -m ilswiss_b200.launch -e exp_specs/sac/sac_hopper.yaml -g 0 --reference /path/to/ILSwiss [--log-dir DIR]
    python -m ilswiss_b200.launch --script run_scripts/sac_alpha_exp_script.py -e variant_0.yaml -g 0 --reference ...

Nothing of the reference is copied or edited: its yaml exp_specs, its `run_scripts/*_exp_script.py` and its `rlkit`
package are used as they are.  What this module does instead of `run_experiment.py` (reference file:line in brackets):

  * variants: `rlkit.launchers.launcher_util.build_nested_variant_generator(exp_specs)` expands the spec exactly as the
    reference launcher does [run_experiment.py:24-45]; the variant yamls are written below the log directory;
  * log directory: `rlkit.launchers.config.LOCAL_LOG_DIR` points INSIDE the reference checkout [config.py:9] -- read-only in
    many deployments -- so it is redirected (`--log-dir`, else $ILSWISS_LOG_DIR, else ./logs) before anything logs
    [used at run_experiment.py:30-35 and launcher_util.py:201-206];
  * one child process per variant, at most meta_data.num_workers at a time [run_experiment.py:50-86] -- but the child is
    `python -m ilswiss_b200.launch --script <meta_data.script_path> ...`, which installs `ilswiss_b200.dropin` and then
    executes the UNMODIFIED experiment script with runpy: the script's own `from rlkit...sac_alpha import SoftActorCritic`
    etc. then bind the device classes (see dropin.py for the table).

A variant can opt out with the optional yaml key `ilswiss_b200: false` (or the flag --pure-reference): the script then runs
exactly as under the reference's launcher.
"""
import argparse
import datetime
import os
import runpy
import subprocess
import sys
import time


def _use_reference(root):
    if root:
        root = os.path.abspath(root)
        if root not in sys.path:
            sys.path.insert(0, root)
    import rlkit  # noqa: F401  (fails loudly when the checkout is not importable)
    return root


def _redirect_logs(log_dir):
    from rlkit.launchers import config

    log_dir = os.path.abspath(log_dir or os.environ.get("ILSWISS_LOG_DIR") or os.path.join(os.getcwd(), "logs"))
    os.makedirs(log_dir, exist_ok=True)
    config.LOCAL_LOG_DIR = log_dir
    return log_dir


def run_script(script, spec, gpu, reference, log_dir, pure):
    """Child side: one variant through the reference's own experiment script."""
    import yaml

    root = _use_reference(reference)
    _redirect_logs(log_dir)
    with open(spec) as f:
        variant = yaml.safe_load(f)
    if not pure and variant.get("ilswiss_b200", True):
        from . import dropin

        dropin.install(root)
    if root and not os.path.isabs(script):
        script = os.path.join(root, script)
    sys.argv = [script, "-e", spec, "-g", str(gpu)]
    runpy.run_path(script, run_name="__main__")


def run_experiment(spec_file, gpu, reference, log_dir, pure):
    """Parent side: expand the exp_spec into variants and run them, num_workers at a time."""
    import yaml

    root = _use_reference(reference)
    log_dir = _redirect_logs(log_dir)
    from rlkit.launchers.launcher_util import build_nested_variant_generator

    with open(spec_file) as f:
        exp_specs = yaml.safe_load(f)
    meta = exp_specs["meta_data"]
    stamp = datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")
    vdir = os.path.join(log_dir, "variants-for-" + meta["exp_name"], "variants-" + stamp)
    os.makedirs(vdir)
    with open(os.path.join(vdir, "exp_spec_definition.yaml"), "w") as f:
        yaml.dump(exp_specs, f, default_flow_style=False)
    specs = []
    for i, variant in enumerate(build_nested_variant_generator(exp_specs)()):
        variant["exp_id"] = i
        path = os.path.join(vdir, "%d.yaml" % i)
        with open(path, "w") as f:
            yaml.dump(variant, f, default_flow_style=False)
        specs.append(path)
    workers = max(1, min(int(meta.get("num_workers", 1)), len(specs)))
    base = [sys.executable, "-m", "ilswiss_b200.launch", "--script", meta["script_path"], "-g", str(gpu), "--log-dir", log_dir]
    if root:
        base += ["--reference", root]
    if pure:
        base += ["--pure-reference"]
    running, todo, failed = [], list(specs), 0
    while todo or running:
        while todo and len(running) < workers:
            cmd = base + ["-e", todo.pop(0)]
            print(" ".join(cmd))
            running.append(subprocess.Popen(cmd))
        time.sleep(0.5)
        still = []
        for p in running:
            rc = p.poll()
            if rc is None:
                still.append(p)
            elif rc != 0:
                failed += 1
        running = still
    return failed


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("-e", "--experiment", required=True, help="exp_spec yaml (or, with --script, one variant yaml)")
    ap.add_argument("-g", "--gpu", type=int, default=0)
    ap.add_argument("--reference", default=os.environ.get("ILSWISS_REFERENCE_ROOT"), help="ILSwiss checkout (directory containing rlkit/)")
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--script", default=None, help="run ONE variant through this experiment script (child mode)")
    ap.add_argument("--pure-reference", action="store_true", help="do not install the drop-in: plain reference run")
    a = ap.parse_args(argv)
    if a.script:
        run_script(a.script, a.experiment, a.gpu, a.reference, a.log_dir, a.pure_reference)
        return 0
    return 1 if run_experiment(a.experiment, a.gpu, a.reference, a.log_dir, a.pure_reference) else 0


if __name__ == "__main__":
    sys.exit(main())
