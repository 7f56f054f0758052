"""Loader for the reference's own native modules built by oracle/Makefile into oracle/_ref/ (compressai.ans: the rANS
coder of cpp_exts/rans/rans_interface.cpp; compressai._CXX: pmf_to_quantized_cdf of cpp_exts/ops/ops.cpp).
TEST INFRASTRUCTURE ONLY - imported by tests/ to check the C-ABI coder against the real reference binary; the product
package never imports it. `load()` returns None when the modules have not been built (no /root/reference at hand)."""
import importlib.util
import os
import sysconfig

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _load_one(name: str):
    path = os.path.join(_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """(ans, _CXX) modules of the reference, or None if oracle/_ref has not been built."""
    ans, cxx = _load_one("ans"), _load_one("_CXX")
    if ans is None or cxx is None:
        return None
    return ans, cxx
