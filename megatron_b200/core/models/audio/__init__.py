from .audio_encoder import AudioEncoderModel, get_num_audio_embeddings  # noqa: F401
