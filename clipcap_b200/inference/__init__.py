from clipcap_b200.inference.base import (generate_beam, generate_beam_tokens, generate_greedy_tokens,  # noqa: F401
                                         generate_nucleus_sampling, generate_no_beam)
