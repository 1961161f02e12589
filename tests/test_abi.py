"""include/vpic_b200_abi.h and vpic_b200/abi.py must match the reference's struct layouts (golden values probed
from the reference headers by tests/golden/make_abi_layout.py).  CPU only."""
import ctypes as C
import json
import os
import subprocess
import tempfile
import pytest

from vpic_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "abi_layout.json")))

NAMES = {"particle_t": "vpb_particle_t", "particle_mover_t": "vpb_particle_mover_t",
         "particle_injector_t": "vpb_particle_injector_t", "species_t": "vpb_species_t", "grid_t": "vpb_grid_t",
         "interpolator_t": "vpb_interpolator_t", "interpolator_array_t": "vpb_interpolator_array_t",
         "accumulator_t": "vpb_accumulator_t", "accumulator_array_t": "vpb_accumulator_array_t",
         "field_t": "vpb_field_t", "field_advance_kernels_t": "vpb_field_advance_kernels_t",
         "field_array_t": "vpb_field_array_t", "material_coefficient_t": "vpb_material_coefficient_t",
         "sfa_params_t": "vpb_sfa_params_t", "hydro_t": "vpb_hydro_t", "hydro_array_t": "vpb_hydro_array_t"}


@pytest.mark.parametrize("simd", ["4", "8", "16"])
def test_c_header_layout(simd):
    lines = []
    for key in GOLD[simd]:
        if key.startswith("sizeof("):
            t = key[7:-1]
            lines.append(f'printf("\\"{key}\\": %zu,\\n", sizeof({NAMES[t]}));')
        elif key.startswith("offsetof("):
            t, m = key[9:-1].split(",")
            lines.append(f'printf("\\"{key}\\": %zu,\\n", offsetof({NAMES[t]},{m}));')
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "vpic_b200_abi.h"\nint main(){printf("{\\n");' + \
          "".join(lines) + 'printf("\\"end\\": 0}\\n");return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", f"-DVPB_SIMD_WIDTH={simd}", "-I", os.path.join(HERE, "..", "include"),
                               os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        got = json.loads(subprocess.check_output([os.path.join(d, "p")]))
    assert got == GOLD[simd]


def test_ctypes_layout():
    g = GOLD["4"]
    assert C.sizeof(abi.Particle) == g["sizeof(particle_t)"] == 32
    assert C.sizeof(abi.ParticleMover) == g["sizeof(particle_mover_t)"] == 16
    assert C.sizeof(abi.ParticleInjector) == g["sizeof(particle_injector_t)"] == 48
    assert C.sizeof(abi.Species) == g["sizeof(species_t)"]
    assert C.sizeof(abi.Grid) == g["sizeof(grid_t)"]
    assert C.sizeof(abi.FieldArray) == g["sizeof(field_array_t)"]
    assert C.sizeof(abi.MaterialCoefficient) == g["sizeof(material_coefficient_t)"]
    assert C.sizeof(abi.HydroArray) == g["sizeof(hydro_array_t)"] and abi.HYDRO_FLOATS * 4 == g["sizeof(hydro_t)"] == 64
    assert abi.HydroArray.g.offset == g["offsetof(hydro_array_t,g)"] and abi.HydroArray.stride.offset == g["offsetof(hydro_array_t,stride)"]
    for m in ("q", "np", "p", "nm", "pm", "last_sorted", "sort_interval", "partition", "g", "id", "next"):
        assert getattr(abi.Species, m).offset == g[f"offsetof(species_t,{m})"]
    for m in ("step", "t0", "x0", "nx", "dx", "rdx", "sx", "nv", "bc", "range", "neighbor", "rangel", "rangeh", "mp"):
        assert getattr(abi.Grid, m).offset == g[f"offsetof(grid_t,{m})"]
    assert abi.particle_dtype.itemsize == 32 and abi.mover_dtype.itemsize == 16 and abi.injector_dtype.itemsize == 48
    for w in (4, 8, 16):
        assert abi.interpolator_floats(w) * 4 == GOLD[str(w)]["sizeof(interpolator_t)"]
        assert abi.accumulator_floats(w) * 4 == GOLD[str(w)]["sizeof(accumulator_t)"]
    assert abi.FIELD_FLOATS * 4 == g["sizeof(field_t)"]
