"""CPU oracle for the rnn-speech acoustic-model path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.

Pinning status (see DESIGN.md "Oracle"):
  * features.fbank      -- PINNED against the reference's own
                           util/audioprocessor.py::_extract_fbank executed in
                           this container (oracle/ref_shim.py +
                           oracle/gen_golden.py -> tests/golden/fbank_*.npz).
  * features.delta      -- 'interp' mode is scipy.signal.savgol_filter, the
                           function librosa>=0.6.1 calls (scipy is present, so
                           the third-party body itself runs); 'edge' mode
                           (librosa<=0.6.0) is restated from memory: unpinned.
  * features.mfcc       -- librosa is absent: restated from its published
                           algorithm, PARITY UNPINNED.
  * resample            -- librosa.load's post-decode half (resampy 'kaiser_best'):
                           third-party, absent, restated; cross-checked against
                           torchaudio's Kaiser-sinc resampler and signal
                           properties, PARITY UNPINNED.
  * lstm / ctc / optim  -- TensorFlow 1.x is absent: restated, PARITY UNPINNED
                           upstream; cross-checked here against torch-CPU
                           autograd (LSTM grads) and torch ctc_loss (labels
                           without the blank id), see tests/test_oracle_*.py.
  * labels              -- PINNED by the reference's known-answer tests
                           (util/test_dataProcessor.py:132-229).
"""
