from evavos_b200.networks import (FeatureFusionBlock, KeyEncoder, KeyProjection, ResBlock, UpsampleBlock,  # noqa: F401
                                  ValueEncoder)
