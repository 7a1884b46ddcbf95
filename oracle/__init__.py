"""CPU oracle for the ReGAT implicit-relation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (tf_vqa_regat_b200/)
may import this package; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and there only as the checker or the
reported CPU baseline.

Pinning status:
  * stage 1 (position_emb.py)  -- PINNED: checked against outputs of the
    reference's own /root/reference/model/position_emb.py, committed as
    tests/golden/stage1_*.npz by oracle/make_golden.py.
  * stages 2-3, loss, clip, Adamax -- PINNED TO THE REFERENCE'S OWN CODE, with
    TensorFlow's primitives restated: the reference needs TensorFlow, which is
    not installable here, and ships no tests or golden vectors.
    oracle/make_golden_ref.py therefore executes the reference's unmodified
    model/*.py and train.py (forward, GradientTape, train(), evaluate()) on
    top of oracle/tf_shim -- a stand-in that restates the ~40 TensorFlow/Keras
    primitives those files call -- and commits the outputs as
    tests/golden/refexec_*.npz.  The restatements below agree with them to
    1e-10 (tests/test_refexec.py).  What stays an assumption is that the
    stand-in's primitives do what TensorFlow's documentation says; they are
    checked against independent implementations, not against TensorFlow.
  * question front-end (language_model.py) -- oracle/language_model.py,
    pinned the same way by tests/golden/refexec_question_*.npz
    (oracle/make_golden_ref_question.py), whole model tokens -> logits included.
  * batch assembly (dataset.py:270-355) -- oracle/dataset_collate.py, pinned by
    tests/golden/refexec_collate.npz: the reference's own dataset.py (tensorize /
    split_entries / trim_collate) executed on an in-memory store
    (oracle/make_golden_ref_collate.py; Keras' pad_sequences restated in the stand-in).
  * explicit relation encoders (relation_encoder.py:95-143, position_emb.py:23-90) --
    oracle/explicit_relation.py, pinned by tests/golden/refexec_explicit_*.npz and explicit_build_graph.npz
    (oracle/make_golden_ref_explicit.py), two cases at num_steps 2 and 3.
  * ImplicitRelationEncoder(num_steps > 1) (relation_encoder.py:82-91) -- regat_torch.encoder(num_steps=k),
    pinned by tests/golden/encsteps_*.npz (oracle/make_golden_ref_steps.py).
  * learning-rate schedule / tokenizer -- tests/golden/refexec_lr_schedule.json, refexec_tokens.json
    (oracle/make_golden_ref_schedule.py, oracle/make_golden_ref_tokens.py).
"""
