"""Static-batch KV-cache bookkeeping (reference ``inference_params.py``)."""


class InferenceParams:
    def __init__(self, max_batch_size, max_sequence_length):
        self.max_sequence_length = max_sequence_length
        self.max_batch_size = max_batch_size
        self.current_batch_size = max_batch_size
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.decode_mode = False
        self.key_value_memory_dict = {}

    def swap_key_value_dict(self, batch_idx):
        if not self.key_value_memory_dict:
            raise ValueError("should not swap when dict is empty")
        for layer, (k, v) in self.key_value_memory_dict.items():
            assert len(batch_idx) == k.shape[1]
            self.key_value_memory_dict[layer] = (k[:, batch_idx], v[:, batch_idx])

    def enable_prefill_mode(self):
        self.decode_mode = False

    def enable_decode_mode(self):
        self.decode_mode = True

    def reset(self):
        self.current_batch_size = self.max_batch_size
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.decode_mode = False
        self.key_value_memory_dict.clear()
