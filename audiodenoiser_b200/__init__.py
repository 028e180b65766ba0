"""B200-native (sm_100a) hot path of jimonld2000/AudioDenoiser: waveform -> |STFT| -> UNet -> iSTFT / overlap-add.

Drop-in modules (same names, constants and signatures as the reference's ``code/*.py`` hot-path functions):

    create_train_dataset.audio_to_magnitude_spectrogram     (reference create_train_dataset.py:162-174)
    create_test_dataset.audio_to_spectrogram                (reference create_test_dataset.py:35-41)
    test.griffin_lim_reconstruction                         (reference test.py:29-48)
    model.UNet                                              (reference model.py:53-94)
    data_loader.SpectrogramDataset                          (reference data_loader.py:7-72)

Batched device-resident path: ``spectral.stft_mag_batched`` / ``spectral.istft_batched``, ``pipeline.Denoiser``,
``sharding.ShardedDenoiser``.  Every numeric step runs in hand-written CUDA kernels of ``libadn_b200.so``
(C ABI: include/adn_b200.h); there is no CPU fallback -- without the library or a B200 the calls raise.
"""

__all__ = ["_lib", "spectral", "model", "pipeline", "sharding", "checkpoint", "synth",
           "create_train_dataset", "create_test_dataset", "test", "data_loader"]
__version__ = "0.1.0"
