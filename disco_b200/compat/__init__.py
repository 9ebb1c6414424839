"""Reference-signature adapters (drop-in names of ``disco_theque``), all backed by the CUDA library.

    disco_theque.math_utils.my_stft / my_istft                 -> compat.math_utils
    disco_theque.sigproc_utils.tf_mask / vad_oracle_batch      -> compat.sigproc_utils
    disco_theque.dnn.utils.tf_mask                             -> compat.sigproc_utils.tf_mask
    disco_theque.se_utils.internal_formulas.intern_filter ...  -> compat.internal_formulas
    disco_theque.speech_enhancement.tango.*                    -> compat.tango
    disco_theque.metrics.snr / sd / fw_snr / fw_sd / si_sdr     -> compat.metrics

NumPy arrays (or torch tensors) in, NumPy arrays out, reference shapes and layouts ((F, T) spectra).
"""
