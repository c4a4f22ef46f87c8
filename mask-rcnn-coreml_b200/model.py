"""placeholder (filled in with the dense model)"""
Detection = MaskRCNN = MaskRCNNConfig = None
