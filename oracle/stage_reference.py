"""ORACLE support (test infrastructure): copy the reference's hot-path Python files (180 KB) from /root/reference to
baseline/_ref/ so that the UNMODIFIED reference can be run on the GPU box (gpurun ships git-ignored files;
/root/reference does not exist there).  baseline/_ref/ is git-ignored: reference sources never enter the history."""
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")


def stage():
    if not os.path.isdir(SRC):
        return False
    for sub, pat in (("rectified_spaattn", ".py"), ("utils", "jenga_gilbert.py")):
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for f in os.listdir(os.path.join(SRC, sub)):
            if f.endswith(pat):
                shutil.copyfile(os.path.join(SRC, sub, f), os.path.join(DST, sub, f))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here")
