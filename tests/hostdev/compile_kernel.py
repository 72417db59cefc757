"""TEST DOUBLE -- the "NVRTC" of tests/hostdev/fake_cuda.c: compile generated CUDA source for the host.

usage: compile_kernel.py <kernel.cu> <out.so>
The source is what the bridge hands to the backend (descriptor line + CUDA C++); the result runs the kernel on the
cooperative emulator of tests/cuda_emulation.py.  Exit code 0 on success; compiler messages go to stdout/stderr.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))

from tests import cuda_emulation as emu  # noqa: E402


def main():
    src, out = Path(sys.argv[1]), Path(sys.argv[2])
    try:
        emu.build_param_launcher(src.read_text(), out)
    except Exception as e:  # noqa: BLE001 -- every failure is a "compilation error" of the double
        print(f"hostdev compile failed: {e}")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
