"""The eleven ``--block-*`` command-line flags (reference core/argparser.py:1-12).  ``vars(args)``
of a parser extended with them is the ``settings`` dict every other entry point consumes."""

# (flag, kwargs) -- names, types, defaults and choices are the contract with existing run scripts
_BLOCK_FLAGS = (
    ("--block-policy", dict(type=str, default="rl_semseg",
                            choices=["static", "all", "none", "random", "rl_semseg", "rl_objectdetection"],
                            help="policy name")),
    ("--block-num-classes", dict(type=int, default=19, help="number of output classes of the main task")),
    ("--block-optim-lr", dict(type=float, default=0.0001, help="policy learning rate")),
    ("--block-optim-wd", dict(type=float, default=0.001, help="policy weight decay")),
    ("--block-optim-momentum", dict(type=float, default=0, help="policy optimizer momentum")),
    ("--block-target", dict(type=float, default=0.50, help="target execution percentage")),
    ("--block-complexity-weight", dict(type=float, default=5,
                                       help="weight gamma, setting importance of complexity reward")),
    ("--block-size", dict(type=int, default=128, help="size of blocks in px")),
    ("--block-train-interval", dict(type=int, default=4, help="optimize the policy every N frames")),
    ("--block-cost-momentum", dict(type=float, default=0.9, help="cost momentum")),
    ("--block-policy-verbose", dict(action="store_true", help="print debug info for policy training")),
)


def add_argparser_arguments(parser):
    for flag, kw in _BLOCK_FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def default_settings(**overrides) -> dict:
    """Settings dict with the flags' defaults (handy for programmatic use; not in the reference)."""
    import argparse

    s = vars(add_argparser_arguments(argparse.ArgumentParser()).parse_args([]))
    s.update(overrides)
    return s
