"""TEST INFRASTRUCTURE recipe: stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (git-ignored; it
travels to the GPU box with the gpurun snapshot, where /root/reference does not exist).

    python oracle/make_ref.py [/root/reference]

Copies, byte for byte and with their relative paths (no file is edited):
    h36m/mlp_mixer.py, h36m/conv_mixer_model.py, conv_mixer/encoding/positional_encoder.py   -- the models
    h36m/utils/utils_mixer.py (mpjpe_error) and what it imports: utils/data_utils.py, utils/forward_kinematics.py
    h36m/train_mixer_h36m.py (the training script whose train() the drop-in test drives unchanged) and its imports
Used only by tests/ (checkpoint round trips through the real modules) and by bench.py's CPU legs (``--impl reference``,
``cpu_baseline`` with kind "reference").  Never imported by the product package.
"""
import os
import shutil
import sys

FILES = ["h36m/mlp_mixer.py", "h36m/conv_mixer_model.py", "conv_mixer/encoding/positional_encoder.py",
         "h36m/utils/utils_mixer.py", "utils/data_utils.py", "utils/forward_kinematics.py"]
# the reference's own training script and what it imports: tests/ref_train.py drives its train() UNCHANGED with the drop-in
# modules; optional for everything else
TRAIN_FILES = ["h36m/train_mixer_h36m.py", "h36m/datasets/dataset_h36m.py", "h36m/datasets/dataset_h36m_ang.py",
               "h36m/utils/data_utils.py", "h36m/utils/forward_kinematics.py", "conv_mixer/utils/visualization_helpers_h3m.py"]
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    for f in FILES + TRAIN_FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), d)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified copies of %s from %s (AlekseiZhuravlev/MotionMixerConv), staged by oracle/make_ref.py.\n" % (", ".join(FILES), src))
    return True


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def train_script_available():
    return available() and all(os.path.exists(os.path.join(DST, f)) for f in TRAIN_FILES)


def import_reference():
    """-> (MlpMixer, ConvMixer, mpjpe_error) of the real reference, imported from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is not staged (run `python oracle/make_ref.py` where /root/reference exists)")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from h36m.mlp_mixer import MlpMixer
    from h36m.conv_mixer_model import ConvMixer
    from h36m.utils.utils_mixer import mpjpe_error
    return MlpMixer, ConvMixer, mpjpe_error


if __name__ == "__main__":
    ok = make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged" if ok else "reference tree not found", DST)
