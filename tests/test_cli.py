"""CLI surface: sub-commands and flags of zernike_master.py:2917-3216 stay unchanged."""
import numpy as np
import pytest

import zernike_master as zm_cli

DEC_FLAGS = ["--mode", "--input", "--order", "--apix", "--com", "--crop_to", "--bin", "--compute_mode", "--gpu_id",
             "--cpu_tasks", "--output", "--bit8", "--save_bit8", "--parity", "--mixed", "--onthefly", "--save",
             "--verbose", "--preallocate_memory_off"]
REC_FLAGS = ["--moments", "--order", "--mode", "--mixed", "--dimensions", "--z_ind", "--scales", "--bit8", "--parity",
             "--apix", "--output", "--cpu_tasks", "--save", "--gpu_id", "--compute_mode", "--verbose",
             "--preallocate_memory_off"]


def _flags(parser, sub):
    sp = [a for a in parser._actions if a.dest == "operating_mode"][0].choices[sub]
    return {s for a in sp._actions for s in a.option_strings}


def test_flags_unchanged():
    p = zm_cli.build_parser()
    assert set(DEC_FLAGS) <= _flags(p, "decomposition")
    assert set(REC_FLAGS) <= _flags(p, "reconstruct")


def test_readme_command_lines_parse():
    p = zm_cli.build_parser()
    a = p.parse_args("decomposition --mode spharm --input meprin_64pix.mrc --order 50 --compute_mode 3 "
                     "--output test --gpu_id 0 --verbose".split())
    assert a.output == "test.npz" and a.order == 50 and a.gpu_id == 0 and a.verbose  # .npz appended (zm:2947-2951)
    a = p.parse_args("reconstruct --mode spharm --moments test.npz --scales -34 0 0 0 --order 50 --compute_mode 3 "
                     "--output reconstructed.mrc --gpu_id 0 --verbose --dimensions 64".split())
    assert a.scales == ["-34", "0", "0", "0"] and a.dimensions == 64 and a.z_ind is None
    a = p.parse_args("reconstruct --mode spharm --moments test.npz --z_ind 3450 --order 50 --output r.mrc".split())
    assert a.z_ind == 3450


@pytest.mark.parametrize("argv", [
    "decomposition --mode spharm --input a.txt --order 5 --output o",             # bad extension
    "decomposition --mode spharm --input a.mrc --output o",                        # --order required
    "reconstruct --mode spharm --moments a.npz --order 5 --output o.npz",          # output must be .mrc
    "decomposition --mode spharm --input a.mrc --order 5 --output o --com c.cmm",  # --com needs --crop_to
    "decomposition --mode monomial --input a.mrc --order 5 --output o",            # monomial cannot run in the reference either
])
def test_invalid_command_lines_exit_2(argv):
    with pytest.raises(SystemExit) as e:
        zm_cli.main(argv.split())
    assert e.value.code == 2


def test_marker_and_crop_helpers():
    from zernike3d_b200.spherical import ZernikeSpherical as Z

    xml = '<marker_set name="m"><marker id="1" x="28.0" y="42.0" z="14.0" r="1" g="1" b="0" radius="1.4"/></marker_set>'
    assert Z.parse_marker_set(xml) == (10, 30, 20)  # (z, y, x) / radius, truncated (zm:2817-2836)
    v = np.arange(20 ** 3, dtype=np.float32).reshape(20, 20, 20)
    c = Z.bin_crop_array_center(v, 8, bin=None, center=(10, 9, 8))
    assert c.shape == (8, 8, 8) and c[0, 0, 0] == v[6, 5, 4]
    b = Z.bin_crop_array_center(v, 9, bin=2, center=(10, 10, 10))  # odd crop rounds down to 8, then /2
    assert b.shape == (4, 4, 4)
    assert [x.shape for x in Z.bin_crop_array_center([v, v], 8, center=(10, 10, 10))] == [(8, 8, 8)] * 2
