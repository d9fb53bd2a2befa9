"""The C++17 host layer (iga_ads_b200/include/ads/*.hpp: the reference's class surface on top of the C ABI)
and the reference's examples rebuilt against it (iga_ads_b200/examples)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "iga_ads_b200", "examples")
PROGS = ("heat_3d", "heat_2d", "implicit_2d", "scalability_3d", "surface_check", "heat_3d_slabs", "element_loop_check", "implicit_3d", "scalability_2d", "generalised_ads")


def build():
    subprocess.run(["make", "-s", "-C", EX], check=True)


def run(prog, *args):
    return subprocess.run([os.path.join(EX, "build", prog), *map(str, args)], capture_output=True, text=True, timeout=600)


def checksum(out):
    return float(re.search(r"sum\(u\) = (-?[0-9.eE+-]+)", out).group(1))


def test_examples_build_with_plain_cxx17():
    build()
    for p in PROGS:
        assert os.path.exists(os.path.join(EX, "build", p))


def test_examples_fail_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    build()
    r = run("heat_3d", 4, 1)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_heat_3d_example_reproduces_the_reference_checksum():
    """BASELINE.json configs[0] through the C++ surface: heat_3d p=2, 12^3, dt=1e-7, 100 steps from the
    shipped initial state; checksum of the compiled reference."""
    build()
    r = run("heat_3d", 12, 100)
    assert r.returncode == 0, r.stderr
    assert abs(checksum(r.stdout) - 132.96044839648852) < 1e-8
    norm = float(re.search(r"\|u\|_2 = ([0-9.]+)", r.stdout).group(1))
    assert abs(norm - 10.367682819844902) < 1e-9
    rq = run("heat_3d", 12, 100, 1)  # same run with the general quadrature kernel
    assert abs(checksum(rq.stdout) - 132.96044839648852) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("prog,args,name,p,ne,dt,steps", [
    ("heat_2d", (3, 24, 5), "heat_2d", 3, 24, 1e-5, 5),
    ("implicit_2d", (3, 24, 5, 1e-2), "implicit_2d", 3, 24, 1e-2, 5),
    ("scalability_3d", (2, 8, 2), "scalability_3d", 2, 8, 1e-6, 2),
])
def test_examples_match_reference_golden(golden, prog, args, name, p, ne, dt, steps):
    """shipped initial state + steps, against the compiled reference's output for the same run"""
    build()
    r = run(prog, *args)
    assert r.returncode == 0, r.stderr
    want = golden["problems"][f"{name}_p{p}_n{ne}_shipped"]
    assert abs(checksum(r.stdout) - want.sum()) < 1e-9 * max(1.0, np.abs(want).sum())


@pytest.mark.gpu
def test_surface_check_value_semantics_projection_norms_sampling():
    """lin::tensor copies taken from a device-resident tensor, buffer-id recycling over 3 x ADSB_MAX_BUFFERS
    temporaries, projection of an arbitrary host callable on the device, errorL2 / normL2 / normH1, sample()"""
    build()
    r = run("surface_check", 12)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "surface_check OK" in r.stdout


@pytest.mark.gpu
# more than two ranks need slabs thick enough for the z factor to couple neighbours only (~47 rows at p = 2)
@pytest.mark.parametrize("n,steps,ranks,p", [(12, 100, 2, 2), (200, 2, 4, 2), (150, 3, 3, 2), (96, 2, 2, 3)])
def test_cxx_slab_host_matches_the_single_gpu_run(n, steps, ranks, p):
    """the z-slab sharded step driven from one C++17 process (adsb_slabs_*: ranks on one device each when the box has
    them, else virtual ranks on device 0; fused distributed z sweep, event barriers) against the single-GPU example,
    and for the BASELINE configs[0] run against the reference's checksum"""
    build()
    r = run("heat_3d_slabs", n, steps, ranks, p)
    assert r.returncode == 0, r.stderr + r.stdout
    got_sum = checksum(r.stdout)
    got_norm = float(re.search(r"\|u\|_2 = ([0-9.]+)", r.stdout).group(1))
    if (n, steps, p) == (12, 100, 2):
        assert abs(got_sum - 132.96044839648852) < 1e-8
        assert abs(got_norm - 10.367682819844902) < 1e-9
    one = run("heat_3d_slabs", n, steps, 1, p)
    assert one.returncode == 0, one.stderr + one.stdout
    assert abs(got_sum - checksum(one.stdout)) < 1e-10 * max(1.0, abs(got_sum))
    want_norm = float(re.search(r"\|u\|_2 = ([0-9.]+)", one.stdout).group(1))
    assert abs(got_norm - want_norm) < 1e-11 * want_norm


@pytest.mark.parametrize("p,ne,threads", [(2, 6, 1), (3, 5, 3), (2, 9, 4)])
def test_reference_style_host_element_loop_runs_unchanged(oracle, p, ne, threads):
    """no GPU: a compute_rhs() in the idiom of examples/scalability/test3d.hpp:66-95 (galois_executor::for_each,
    element_rhs, eval_fun / eval_basis / grad_dot, synchronized + update_global_rhs, tensor_view) compiled against
    the C++17 headers reproduces the oracle's right-hand side of the same problem"""
    build()
    r = run("element_loop_check", p, ne, threads)
    assert r.returncode == 0, r.stderr + r.stdout
    got_sum = float(re.search(r"sum\(rhs\) = (-?[0-9.eE+-]+)", r.stdout).group(1))
    got_norm = float(re.search(r"\|rhs\|_2 = ([0-9.eE+-]+)", r.stdout).group(1))
    n = ne + p
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    u_prev = (np.sin(0.3 * i) + 0.5 * np.cos(0.2 * j) + 0.1 * k).ravel(order="F")
    want, _ = oracle.run("scalability_3d", p, ne, 1e-6, 1, u0=u_prev, stage=1)
    assert abs(got_sum - want.sum()) < 1e-12 * np.abs(want).sum()
    assert abs(got_norm - np.linalg.norm(want)) < 1e-12 * np.linalg.norm(want)
