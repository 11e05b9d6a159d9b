from evavos_b200.inference_core import InferenceCore  # noqa: F401
