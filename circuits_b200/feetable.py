"""Fee computation of the Hermez protocol (host mirror of commonjs
`feeTable.computeFee` [un-vendored]; circuit: /root/reference/src/compute-fee.circom:12-94)."""
from .fee_table import FEE_SHIFT_TABLE


def compute_fee(amount, sel):
    if sel < 192:
        return (amount * FEE_SHIFT_TABLE[sel]) >> 60
    return amount * FEE_SHIFT_TABLE[sel]
