"""Shared helpers of the parity tests."""
import torch


def check_tokens_against_oracle(tokens, lengths, oracle_results, margin_tol):
    """Greedy token ids must equal the oracle's bit for bit. fp16 operands give logit errors ~1e-3 * |logit|max, so a
    mismatch is only tolerated at a step where the ORACLE's own top-1/top-2 margin is below `margin_tol` (a near-tie the
    reference itself would flip under reordering of its fp32 sums); that sample is then no longer compared past the
    flip. Returns (n_exact_samples, n_flips)."""
    tokens, lengths = tokens.cpu().tolist(), lengths.cpu().tolist()
    exact = flips = 0
    for i, (otoks, _score, trace) in enumerate(oracle_results):
        diverged = False
        for s, ot in enumerate(otoks):
            if tokens[i][s] != ot:
                margin = trace[s]["margin"][0]
                assert margin < margin_tol, (
                    f"sample {i} step {s}: got {tokens[i][s]}, oracle {ot}, oracle margin {margin:.3e} >= {margin_tol:.1e}")
                diverged = True
                flips += 1
                break
        if not diverged:
            assert lengths[i] == len(otoks), f"sample {i}: length {lengths[i]} != oracle {len(otoks)}"
            exact += 1
    return exact, flips
