"""Generates tests/golden/abi_layout.json from the REFERENCE headers (run in the build container only:
needs /root/reference and g++).  The test suite compares include/vpic_b200_abi.h and vpic_b200/abi.py against it."""
import json, os, subprocess, sys, tempfile

REF = os.environ.get("VPIC_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "..", "..", "oracle", "mpi_shim")

PROBE = r'''
#include <cstddef>
#include <cstdio>
#include "species_advance/species_advance.h"
#include "sf_interface/sf_interface.h"
#define IN_sfa
#include "field_advance/standard/sfa_private.h"
#define S(T) printf("\"sizeof(%s)\": %zu,\n", #T, sizeof(T))
#define O(T,m) printf("\"offsetof(%s,%s)\": %zu,\n", #T, #m, offsetof(T,m))
int main() {
  printf("{\n");
  S(particle_t); O(particle_t,i); O(particle_t,ux); O(particle_t,w);
  S(particle_mover_t); O(particle_mover_t,i);
  S(particle_injector_t); O(particle_injector_t,dispx); O(particle_injector_t,sp_id);
  S(species_t); O(species_t,q); O(species_t,np); O(species_t,p); O(species_t,nm); O(species_t,pm);
  O(species_t,last_sorted); O(species_t,sort_interval); O(species_t,partition); O(species_t,g); O(species_t,id); O(species_t,next);
  S(grid_t); O(grid_t,step); O(grid_t,t0); O(grid_t,x0); O(grid_t,nx); O(grid_t,dx); O(grid_t,rdx); O(grid_t,sx);
  O(grid_t,nv); O(grid_t,bc); O(grid_t,range); O(grid_t,neighbor); O(grid_t,rangel); O(grid_t,rangeh); O(grid_t,mp);
  S(interpolator_t); O(interpolator_t,cbx); O(interpolator_t,dcbzdz);
  S(interpolator_array_t); O(interpolator_array_t,g);
  S(accumulator_t); O(accumulator_t,jy); O(accumulator_t,jz);
  S(accumulator_array_t); O(accumulator_array_t,n_pipeline); O(accumulator_array_t,stride); O(accumulator_array_t,g);
  S(field_t); O(field_t,cbx); O(field_t,tcax); O(field_t,jfx); O(field_t,ematx); O(field_t,cmat);
  S(field_advance_kernels_t); S(field_array_t); O(field_array_t,g); O(field_array_t,params); O(field_array_t,kernel);
  S(material_coefficient_t); S(sfa_params_t); O(sfa_params_t,n_mc); O(sfa_params_t,damp);
  S(hydro_t); O(hydro_t,rho); O(hydro_t,ke); O(hydro_t,tyz); O(hydro_t,txy);
  S(hydro_array_t); O(hydro_array_t,n_pipeline); O(hydro_array_t,stride); O(hydro_array_t,g);
  printf("\"end\": 0\n}\n");
  return 0;
}
'''

def probe(flags):
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "probe.cc"); exe = os.path.join(d, "probe")
        open(src, "w").write(PROBE)
        subprocess.check_call(["g++", "-std=c++11", "-w", "-DVPIC_USE_PTHREADS", f"-I{SHIM}", f"-I{REF}/src", f"-I{REF}"] + flags + [src, "-o", exe])
        return json.loads(subprocess.check_output([exe]))

out = {"4": probe([]), "8": probe(["-mavx2", "-mfma", "-DUSE_V4_AVX2", "-DUSE_V8_AVX2"]),
       "16": probe(["-mavx2", "-mfma", "-DUSE_V4_AVX2", "-DUSE_V16_PORTABLE"])}
json.dump(out, open(os.path.join(HERE, "abi_layout.json"), "w"), indent=1, sort_keys=True)
print("wrote abi_layout.json")
