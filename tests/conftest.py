import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def core_lib():
    from lighthouse2_b200 import load_library
    return load_library()


# Order of the GPU suite: the parity tests proper first (ray queries, frames, the shade stage against the reference's kernel, the
# filter chain), then the features built on them, and the tests that spawn other processes (hosts, torchrun) last - so that a failure
# in a peripheral test can never hide the BASELINE-size parity tests behind `-x`.
_ORDER = ["test_traversal_gpu", "test_render_gpu", "test_shade_stage_gpu", "test_filter_gpu", "test_filter_mode_gpu",
          "test_call_contract_gpu", "test_animation_gpu", "test_tile_gpu", "test_frame_structure_gpu", "test_ref_timing_gpu",
          "test_multigpu_gpu", "test_rendersystem_dropin", "test_dropin_gpu"]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _ORDER.index(name) if name in _ORDER else -1     # CPU-side files keep their place in front
    items.sort(key=key)                                         # stable: order within a file is untouched
