"""Token ids the reference's own dataset.py produces for a handful of questions: Dictionary.tokenize (dataset.py:63-77) and
VQAFeatureDataset.tokenize (:250-263), executed over the stand-ins of oracle/tf_shim.  Build container only:

    python -m oracle.make_golden_ref_tokens        # rewrites tests/golden/refexec_tokens.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

QUESTIONS = ["What is the man's hat color?", "Is this a kitchen, or a bathroom?", "How many zebras are there?",
             "what's on the table", "Is   the  dog's  tail wagging ?", "Why", "",
             "What color is the very long striped scarf that the tall woman standing next to the red double decker bus is wearing today?",
             "Are the people's umbrellas open, closed, or broken?"]
WORDS = ["what", "is", "the", "man", "'s", "hat", "color", "this", "a", "kitchen", "or", "bathroom", "how", "many", "are", "there", "on",
         "table", "dog", "tail", "why", "very", "long", "scarf", "that", "tall", "woman", "standing", "next", "to", "red", "bus", "wearing",
         "people", "open", "closed", "bebe"]


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, "/root/reference")
    import dataset as ref_dataset
    d = ref_dataset.Dictionary()
    for w in WORDS:
        d.add_word(w)
    ds = object.__new__(ref_dataset.VQAFeatureDataset)
    ds.dictionary = d
    ds.entries = [{"question": q} for q in QUESTIONS]
    ds.tokenize()                                                     # default max_length = 14
    out = {"word2idx": d.word2idx, "ntoken": d.ntoken, "padding_idx": d.padding_idx, "questions": QUESTIONS,
           "q_token": [list(map(int, e["q_token"])) for e in ds.entries]}
    with open(os.path.join(ROOT, "tests", "golden", "refexec_tokens.json"), "w") as f:
        json.dump(out, f, indent=1)
    for q, e in zip(QUESTIONS, ds.entries):
        print(e["q_token"], q[:40])


if __name__ == "__main__":
    main()
