"""The C++ host layer (include/heros_host.hpp: the reference's plug-in classes -- Covid19TransmissionArea / Distance,
Covid19Progression, DiseasePolicy, LocationPolicy, HerosModel -- over the C ABI) through its own test program
tests/cpp/test_host.cpp.  The reference is Java and there is no JVM here, so this compiled host stands where the Java
classes would; the program checks itself against the oracle and returns non-zero on any failed check."""
import os
import subprocess

import pytest

from common import GOLDEN, ROOT

CPP = os.path.join(ROOT, "tests", "cpp")


def run(mode):
    subprocess.check_call(["make", "-C", CPP], stdout=subprocess.DEVNULL)
    env = dict(os.environ)
    extra = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime
        extra.insert(0, os.path.join(list(nvidia.cuda_runtime.__path__)[0], "lib"))
    except ImportError:
        pass
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([os.path.join(CPP, "test_host"), mode, GOLDEN], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr)
    return r


def test_cpp_host_logic_without_a_device():
    """.properties reader, ConditionalProbability / distribution syntax, MedlabsException on bad input, and no CPU
    fallback: constructing the model without a usable GPU fails with HEROS_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal to start cannot be observed (covered by the gpu mode)")
    r = run("cpu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_host_against_the_oracle_on_the_device():
    """Javadoc worked example and boundary cases through DiseaseTransmission.infectPeople, DiseasePolicy /
    LocationPolicy error behaviour, six-day windowed runs of both models equal to the oracle's event lists."""
    r = run("gpu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout and r.stdout.count("identical to the oracle") == 3
