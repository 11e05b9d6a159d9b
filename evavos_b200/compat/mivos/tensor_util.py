from evavos_b200.tensor_util import pad_divide_by, unpad, unpad_3dim  # noqa: F401
