from evavos_b200.networks import FusionNet  # noqa: F401
