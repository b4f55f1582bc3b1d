from .engine import DynamicInferenceEngine, InferenceRequest, StaticInferenceEngine
from .kv_cache import KVBlockAllocator, PagedKVCache
from .sampling import SamplingParams, sample
