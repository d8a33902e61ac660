"""not-gpu: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares,
and fails loudly (no CPU fallback) when no device is usable."""
import ctypes as C
import subprocess

import pytest
import torch

from multimodal_drl_rmc_b200 import _lib


@pytest.fixture(scope="module")
def built():
    return _lib.build_library()


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", built], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    assert [s for s in declared if s not in exported] == []
    assert sorted(_lib._SIGS) == declared, "ctypes signature table must cover the header exactly"


def test_library_loads_and_reports_abi(built):
    h = _lib.lib()
    assert h.rmc_abi_version() == _lib.ABI_VERSION == 3
    assert h.rmc_launch_count() >= 0


def test_sass_contains_tma_bulk_copy_and_mbarrier(built):
    sass = subprocess.check_output(["cuobjdump", "-sass", built], text=True)
    assert "UBLKCP" in sass, "parameter staging must use the TMA bulk-copy engine"
    assert "SYNCS" in sass
    assert "sm_100a" in subprocess.check_output(["cuobjdump", "-lelf", built], text=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(built):
    h = C.c_void_p()
    rc = _lib.lib().rmc_replay_create(C.byref(h), 100, 14, 1, 0)
    assert rc == -2 and b"cuda" in _lib.lib().rmc_last_error().lower()
    from multimodal_drl_rmc_b200 import macro_config
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        macro_config.make_agent("DuelingDoubleDQNAgent", 14, 32, 1000, save_dir="/tmp/rmc_x/", log_dir="/tmp/rmc_x/")
    from multimodal_drl_rmc_b200 import Networks
    net = Networks.DuelingDeepQNetwork(torch.device("cpu"), 1e-4, macro_config.network_config, macro_config.ObsSpace(14), 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.actions([[0.0] * 14])


def test_product_package_never_imports_the_oracle():
    import os
    import re
    pkg = _lib.PKG_DIR
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
