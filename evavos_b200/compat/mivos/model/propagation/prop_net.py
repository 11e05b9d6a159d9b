from evavos_b200.memory_reader import EvalMemoryReader  # noqa: F401
from evavos_b200.networks import AttentionMemory, Decoder, PropagationNetwork  # noqa: F401
