"""Fee computation of the Hermez protocol (host mirror of commonjs
`feeTable.computeFee` [un-vendored]; circuit: /root/reference/src/compute-fee.circom:12-94).

The table itself exists once in the product: `csrc/fee_table.inc`, the data the device kernels include
(verbatim values of `feeShiftTable(i)`, src/compute-fee.circom:105-109); it is parsed from there."""
import os
import re

_INC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "fee_table.inc")
with open(_INC) as _f:
    FEE_SHIFT_TABLE = [int(x) for x in re.findall(r"(\d+)ULL", _f.read())]
assert len(FEE_SHIFT_TABLE) == 256


def compute_fee(amount, sel):
    if sel < 192:
        return (amount * FEE_SHIFT_TABLE[sel]) >> 60
    return amount * FEE_SHIFT_TABLE[sel]
