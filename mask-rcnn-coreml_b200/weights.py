"""placeholder (filled in with the dense model)"""
