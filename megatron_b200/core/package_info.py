MAJOR, MINOR, PATCH = 0, 1, 0
__version__ = f"{MAJOR}.{MINOR}.{PATCH}"
__package_name__ = "megatron_b200"
__description__ = "Blackwell-native Megatron-Core: tensor/sequence/pipeline/expert parallel training for one 8xB200 NVSwitch box"
