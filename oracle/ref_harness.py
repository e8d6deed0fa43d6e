"""Import the *Python reference itself* (read-only, /root/reference) inside this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to generate golden vectors
and by the optional `tests/test_oracle_vs_reference.py` cross-check (skipped when
/root/reference is absent, i.e. on the GPU box).  Nothing in the product package, bench.py's
own arm or smoke() touches this file.

Recipe = SURVEY.md Appendix A: stub modules for xmltodict / pyecharts / tkinter /
mysql.connector (oracle/stubs), a writable scratch CWD with config/, logs/ and
DataBase/experience/ (the reference uses CWD-relative paths everywhere), <save_loop> pushed
far above any run length so the shipped Mod/*.pth are never rewritten (they could not be
anyway: /root/reference is read-only).
"""
import os
import re
import shutil
import sys

REF = os.environ.get("UAVRL_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
WORK = os.path.join(HERE, "_work")


def available():
    return os.path.isdir(os.path.join(REF, "Agents"))


def prepare_workdir(trainer_overrides=None):
    """Create oracle/_work with config/ copied from the reference (save_loop made huge)."""
    os.makedirs(os.path.join(WORK, "logs"), exist_ok=True)
    os.makedirs(os.path.join(WORK, "DataBase", "experience"), exist_ok=True)
    cdir = os.path.join(WORK, "config")
    os.makedirs(cdir, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "config")):
        if f.endswith(".xml"):
            shutil.copy(os.path.join(REF, "config", f), os.path.join(cdir, f))
            os.chmod(os.path.join(cdir, f), 0o644)
    tp = os.path.join(cdir, "Trainer.xml")
    txt = open(tp).read()
    txt = re.sub(r"<save_loop>\d+</save_loop>", "<save_loop>1000000000</save_loop>", txt)
    for k, v in (trainer_overrides or {}).items():
        txt = re.sub(r"<%s>[^<]*</%s>" % (k, k), "<%s>%s</%s>" % (k, v, k), txt)
    open(tp, "w").write(txt)
    return WORK


def load_reference():
    """chdir into the scratch dir, put stubs + reference on sys.path, import `simulator`.

    Returns the imported `simulator` module (importing it seeds random/numpy/torch with 42,
    simulator.py:31-39).
    """
    if not available():
        raise RuntimeError("reference not present at %s" % REF)
    prepare_workdir()
    os.chdir(WORK)
    for p in (os.path.join(HERE, "stubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    import simulator  # noqa: E402  (the reference's launcher)
    return simulator
