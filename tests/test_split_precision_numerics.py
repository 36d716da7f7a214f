"""CPU statement of the numerics behind SNB_PREC_FP32_TC (csrc/mlp_tc.cu, split mode): every MMA operand as two fp16 parts, three
products per product, fp32 accumulation truncated after every K = 16 instruction.  The emulation (tools/experiments/
split_precision_emulation.py) runs the shipped 3 / 1 / 256 decoder with exactly that arithmetic on the CPU; the GPU kernels are held to
the fp32 tolerance by tests/test_gpu_parity.py, this file pins WHY the design choices are what they are."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "experiments"))
import split_precision_emulation as emu  # noqa: E402

N = 384


def test_two_fp16_parts_and_three_products_are_fp32_grade():
    """Round-to-nearest accumulation: outputs and every gradient as close to fp64 as plain fp32 arithmetic is (within 3x)."""
    out = emu.run(N, modes=("fp32", "fp16x2"))
    for k, (_, e64) in out["fp16x2"].items():
        assert e64 < 3e-6 and e64 < 3 * max(out["fp32"][k][1], 2e-7), (k, e64, out["fp32"][k][1])


def test_two_bf16_parts_are_not_enough():
    """... which two bf16 parts (16 mantissa bits) are not: the reason the operands are fp16 pairs with power-of-two scaling."""
    out = emu.run(N, modes=("fp16x2", "bf16x2"))
    assert out["bf16x2"]["rgb"][1] > 4 * out["fp16x2"]["rgb"][1]


def test_truncating_accumulator_prefers_corrections_first():
    """With the tensor core's truncating accumulator the forward error grows; issuing every correction product of a layer before its
    leading products (the accumulator is still ~2^-11 of its final size while they land) wins most of it back, and the forward stays
    inside the 1e-5 tolerance with an order of magnitude to spare either way."""
    inter = emu.run(N, modes=("fp16x2",), rz=True, corr_first=False)["fp16x2"]
    first = emu.run(N, modes=("fp16x2",), rz=True, corr_first=True)["fp16x2"]
    assert first["rgb"][1] < 0.75 * inter["rgb"][1], (first["rgb"], inter["rgb"])
    assert inter["rgb"][1] < 3e-6 and first["sigma"][1] < 1e-6
