"""Stage the UNMODIFIED reference `nano` package where the GPU box can import it (TEST INFRASTRUCTURE).

/root/reference exists only in the build container.  `__graft_entry__.build()` calls `stage()` there: the reference's
timeviper/model/llm/llm_repo/nano/ package (modeling_nano.py, configuration_nano.py, merge_modules/) is copied byte for
byte into baseline/_ref/nano/ -- git-ignored, so no reference source enters the history, but not gpurun-ignored, so it
travels to the GPU box with the snapshot.  Consumers (tests/test_gpu_reference_model.py, the config-1 leg of
`bench.py --impl reference`) import it through `load()` and skip / report "absent" when the directory is not there.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/timeviper/model/llm/llm_repo/nano"
DST_ROOT = os.path.join(HERE, "..", "baseline", "_ref")
DST = os.path.join(DST_ROOT, "nano")


def stage():
    """Copy the package if the reference tree is present; returns True if baseline/_ref/nano is usable afterwards."""
    if os.path.isdir(SRC):
        for root, _, files in os.walk(SRC):
            rel = os.path.relpath(root, SRC)
            out = os.path.join(DST, rel) if rel != "." else DST
            os.makedirs(out, exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    s, d = os.path.join(root, f), os.path.join(out, f)
                    if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                        shutil.copyfile(s, d)
    return available()


def available():
    return os.path.isfile(os.path.join(DST, "modeling_nano.py"))


def load():
    """Import the staged reference: returns (modeling_nano module, NemotronHConfig).  The module hard-imports
    mamba_ssm's Triton rmsnorm_fn (modeling_nano.py:73-77); oracle/_shim supplies a pure-torch stand-in so that the
    import succeeds without the wheel (the product rebinds all six operator names with patch_reference anyway)."""
    if not available():
        raise ImportError("baseline/_ref/nano is absent: run __graft_entry__.build() in the container that has /root/reference")
    for pth in (os.path.join(HERE, "_shim"), os.path.abspath(DST_ROOT)):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    # On a GPU box transformers' is_mamba_2_ssm_available() finds the shim package and then fails to parse its version
    # ('N/A': the shim has no distribution metadata).  The wheels are absent either way, so the harness answers "not
    # available" for the duration of the import -- the reference module then binds None to the operator names
    # (modeling_nano.py:66-71, :82), exactly what it does in the CPU container; patch_reference rebinds them afterwards.
    import transformers.utils.import_utils as iu
    saved = (iu.is_mamba_2_ssm_available, iu.is_causal_conv1d_available)
    iu.is_mamba_2_ssm_available = lambda: False
    iu.is_causal_conv1d_available = lambda: False
    try:
        import nano.modeling_nano as mn
        from nano.configuration_nano import NemotronHConfig
    finally:
        iu.is_mamba_2_ssm_available, iu.is_causal_conv1d_available = saved
    return mn, NemotronHConfig


if __name__ == "__main__":
    print("staged" if stage() else "reference tree absent; nothing staged")
