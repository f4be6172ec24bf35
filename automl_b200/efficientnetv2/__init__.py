"""EfficientNet V1/V2 classifier backbones on the B200 path (reference: efficientnetv2/)."""
