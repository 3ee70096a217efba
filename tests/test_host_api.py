"""CPU-only tests: the C-ABI library loads and exports every symbol declared
in include/improver_b200.h, and the host logic of the plugin surface behaves
like the reference (constructor validation, error messages, helpers)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from improver_b200 import _lib

    header = open(os.path.join(ROOT, "include", "improver_b200.h")).read()
    declared = set(re.findall(r"\b(nbh_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points found in the header"
    lib = ctypes.CDLL(_lib.LIBPATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert set(_lib.EXPORTED) == declared
    assert _lib.load().nbh_version() == 2


def test_struct_layout_matches_header():
    from improver_b200._lib import NbhDesc, NbhThreshold

    assert ctypes.sizeof(NbhThreshold) == 32
    assert ctypes.sizeof(NbhDesc) == 104
    assert NbhDesc.thresholds.offset == 48 and NbhDesc.mask2d.offset == 56
    assert NbhDesc.plane_cells.offset == 72 and NbhDesc.y_offset.offset == 80 and NbhDesc.ext_stats.offset == 96
    from improver_b200._lib import NbhSliceStats
    assert ctypes.sizeof(NbhSliceStats) == 40


def test_struct_layout_agrees_with_the_compiler(tmp_path):
    """sizeof / offsetof of the header's structs as gcc sees them."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "improver_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(NbhDesc), offsetof(NbhDesc, plane_cells),'
                   'offsetof(NbhDesc, y_offset), offsetof(NbhDesc, ext_stats), sizeof(NbhSliceStats), sizeof(NbhThreshold));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [104, 72, 80, 96, 40, 32]


def test_counter_based_generator_is_sharding_independent():
    """workloads.counter_uniform (config 5 input): torch == numpy restatement, and a shard
    reproduces its part of the whole."""
    import torch
    from improver_b200 import workloads as W

    whole = W.counter_uniform(0, 5, (12, 20), 5, "cpu").numpy()
    np.testing.assert_array_equal(whole, W.counter_uniform_host(0, 5, (12, 20), 5))
    part = W.counter_uniform(3, 2, (12, 20), 5, "cpu").numpy()
    np.testing.assert_array_equal(part, whole[3:5])
    assert 0.0 <= whole.min() and whole.max() < 1.0 and abs(whole.mean() - 0.5) < 0.05


def test_no_cpu_fallback_without_cuda():
    import torch
    from improver_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(TypeError, match="CUDA tensor"):
        engine.neighbourhood(torch.zeros(4, 4), 1)


def test_radius_by_lead_time():
    """improver_tests/nbhood/test_init.py."""
    from improver_b200.nbhood import radius_by_lead_time

    assert radius_by_lead_time(["10000"]) == (10000.0, None)
    assert radius_by_lead_time(["10000", "20000"], ["0", "6"]) == ([10000.0, 20000.0], [0, 6])
    with pytest.raises(ValueError, match="Multiple radii"):
        radius_by_lead_time(["1", "2"])
    with pytest.raises(ValueError, match="equal length"):
        radius_by_lead_time(["1", "2"], ["0"])


def test_neighbourhood_processing_init_errors():
    """improver_tests/nbhood/nbhood/test_NeighbourhoodProcessing.py:18-36,
    test_BaseNeighbourhoodProcessing.py."""
    from improver_b200.nbhood.nbhood import BaseNeighbourhoodProcessing, NeighbourhoodProcessing

    with pytest.raises(ValueError, match="is not a valid neighbourhood_method"):
        NeighbourhoodProcessing("oval", 10000)
    with pytest.raises(ValueError, match="weighted_mode can only be used if neighbourhood_method is circular"):
        NeighbourhoodProcessing("square", 10000, weighted_mode=True)
    with pytest.raises(ValueError, match="mismatch in the number of radii"):
        BaseNeighbourhoodProcessing([10000, 20000], lead_times=[1, 2, 3])
    plugin = BaseNeighbourhoodProcessing([10000, 20000, 30000], lead_times=[1, 3, 5])
    np.testing.assert_allclose(plugin._find_radii(cube_lead_times=np.array([2, 3, 4])), [15000, 20000, 25000])


def test_meta_neighbourhood_init_errors():
    """improver_tests/nbhood/nbhood/test_MetaNeighbourhood.py."""
    from improver_b200.nbhood.nbhood import MetaNeighbourhood

    with pytest.raises(RuntimeError, match="weighted_mode cannot be used with"):
        MetaNeighbourhood("percentiles", 10000, weighted_mode=True)
    with pytest.raises(RuntimeError, match="Cannot generate percentiles from complex numbers"):
        MetaNeighbourhood("percentiles", 10000, degrees_as_complex=True)
    with pytest.raises(RuntimeError, match="Cannot process complex numbers with circular"):
        MetaNeighbourhood("probabilities", 10000, neighbourhood_shape="circular", degrees_as_complex=True)


def test_threshold_init_errors_and_bounds():
    """improver_tests/threshold/test_Threshold.py (init cases), threshold.py:158-239."""
    from improver_b200.threshold import BasicThreshold, Threshold

    assert BasicThreshold is Threshold
    with pytest.raises(ValueError, match="mutually exclusive"):
        Threshold(threshold_values=[1.0], threshold_config={"1.0": "None"})
    with pytest.raises(ValueError, match="One of threshold_config or threshold_values"):
        Threshold()
    with pytest.raises(ValueError, match="Invalid fuzzy_factor"):
        Threshold(threshold_values=1.0, fuzzy_factor=1.5)
    with pytest.raises(ValueError, match="threshold == 0"):
        Threshold(threshold_values=0.0, fuzzy_factor=0.5)
    with pytest.raises(ValueError, match="Threshold must be within bounds"):
        Threshold(threshold_config={"0.6": [0.7, 0.8]})
    with pytest.raises(ValueError, match="does not match any known comparison_operator"):
        Threshold(threshold_values=1.0, comparison_operator="!")
    with pytest.raises(ValueError, match="Can only collapse over"):
        Threshold(threshold_values=1.0, collapse_coord="kittens")
    with pytest.raises(ValueError, match="Cannot apply cell methods without collapsing"):
        Threshold(threshold_values=1.0, collapse_cell_methods={"realization": "mean"})
    plugin = Threshold(threshold_values=[0.5, -2.0], fuzzy_factor=0.5)
    assert plugin.fuzzy_bounds == [(0.25, 0.75), (-3.0, -1.0)]
    assert Threshold(threshold_config={"0.6": [0.4, 0.7]}).fuzzy_bounds == [(0.4, 0.7)]


def test_grid_helpers_and_errors():
    """improver/utilities/spatial.py:42-161, nbhood.py:32-52."""
    from improver_b200.nbhood.nbhood import check_radius_against_distance, circular_kernel
    from improver_b200.synthetic_data import set_up_variable_cube
    from improver_b200.utilities.spatial import check_if_grid_is_equal_area, distance_to_number_of_grid_cells

    cube = set_up_variable_cube(np.zeros((16, 16), np.float32))
    assert distance_to_number_of_grid_cells(cube, 10000) == 5
    assert distance_to_number_of_grid_cells(cube, 11999) == 5
    with pytest.raises(ValueError, match="gives zero cell extent"):
        distance_to_number_of_grid_cells(cube, 1000)
    with pytest.raises(ValueError, match="Please specify a positive distance in metres"):
        distance_to_number_of_grid_cells(cube, -3)
    check_if_grid_is_equal_area(cube)
    with pytest.raises(ValueError, match="exceeds max domain distance"):
        check_radius_against_distance(cube, 50000)
    latlon = set_up_variable_cube(np.zeros((16, 16), np.float32), spatial_grid="latlon")
    with pytest.raises(ValueError):
        check_if_grid_is_equal_area(latlon)
    np.testing.assert_array_equal(circular_kernel(1, False), [[0, 1, 0], [1, 1, 1], [0, 1, 0]])


def test_neighbourhood_tools_views_match_oracle():
    """improver_tests/utilities/test_neighbourhood_tools.py (view helpers)."""
    from improver_b200.utilities import neighbourhood_tools as nt
    from oracle import nbhood_oracle as orc

    data = np.arange(30, dtype=np.float32).reshape(5, 6)
    np.testing.assert_array_equal(nt.rolling_window(data, (3, 3)), orc.rolling_window(data, (3, 3)))
    np.testing.assert_array_equal(nt.pad_and_roll(data, (3, 3), mode="mean", stat_length=1),
                                  orc.pad_and_roll(data, (3, 3), mode="mean", stat_length=1))
    with pytest.raises(ValueError, match="Number of dimensions"):
        nt.rolling_window(data[0], (3, 3))
    with pytest.raises(RuntimeError, match="negative or zero"):
        nt.rolling_window(data, (7, 7))
    with pytest.raises(ValueError, match="odd number"):
        nt.boxsum(data, 4)
    with pytest.raises(ValueError, match="integer type"):
        nt.boxsum(data, 3.0)


def test_post_processed_title():
    """improver/__init__.py:75-89."""
    from improver_b200 import PostProcessingPlugin

    class Dummy:
        attributes = {"title": "UKV Model Forecast"}

    PostProcessingPlugin.post_processed_title(Dummy)
    assert Dummy.attributes["title"] == "Post-Processed UKV Model Forecast"
