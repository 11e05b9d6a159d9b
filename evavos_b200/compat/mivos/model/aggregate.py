from evavos_b200.aggregate import aggregate_wbg  # noqa: F401
