"""Rule Set 1 (Doench 2014) weight table for the ORACLE.

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Nothing in the product
package (cropsr_b200/) imports this file.

The reference carries the weights as two dense float64 literals
(/root/reference/CROPSR.py:165-188 first_matrix[120], :190-283
second_matrix[464]) indexed  w1[4*p + c]  and  w2[16*p + 4*c1 + c2]  with
p the 0-based position in the scored 30-mer and the code order A0 T1 C2 G3
(/root/reference/CROPSR.py:300-302).  Here they are restated sparsely from the
published label/weight lists (the same lists appear in the reference's dead
code, /root/reference/cropsr_functions.py:99-121); labels use 1-based
positions.  `check_digests()` pins the dense arrays to the sha256 digests of
the reference's arrays recorded in SURVEY.md Appendix A.
"""
import hashlib

import numpy as np

CODE = {"A": 0, "T": 1, "C": 2, "G": 3}

INTERCEPT = 0.59763615          # /root/reference/CROPSR.py:161
LOW_GC = -0.2026259             # /root/reference/CROPSR.py:162 (always added, :312)
HIGH_GC = -0.1665878            # /root/reference/CROPSR.py:163 (never used)

FIRST_ORDER = {
    "G02": -0.2753771, "A03": -0.3238875, "C03": 0.17212887, "C04": -0.1006662,
    "C05": -0.2018029, "G05": 0.24595663, "A06": 0.03644004, "C06": 0.09837684,
    "C07": -0.7411813, "G07": -0.3932644, "A12": -0.466099, "A15": 0.08537695,
    "C15": -0.013814, "A16": 0.27262051, "T16": -0.2859442, "C16": 0.1190226,
    "A17": 0.09745459, "G17": -0.1755462, "C18": -0.3457955, "G18": -0.6780964,
    "A19": 0.22508903, "C19": -0.5077941, "T20": -0.054307, "G20": -0.4173736,
    "T21": -0.0907126, "G21": 0.37989937, "T22": -0.5305673, "C22": 0.05782332,
    "T23": -0.8770074, "T24": -0.4031022, "C24": -0.8762358, "G24": 0.27891626,
    "A25": -0.0773007, "T25": -0.2216372, "C25": 0.28793562, "T28": 0.11787758,
    "G28": -0.6890167, "C29": -0.1604453, "G30": 0.38634258,
}

SECOND_ORDER = {
    "GT02": -0.6257787, "GC05": 0.30004332, "AA06": -0.8348362, "TA06": 0.76062777,
    "GG07": -0.4908167, "TA12": 0.7092612, "TT12": -0.5868739, "TC12": 0.49629861,
    "GG12": -1.5169074, "GG13": -0.3345637, "GA14": 0.76384993, "GC14": -0.5370252,
    "TG17": -0.7981461, "TC19": 0.35318325, "GG19": -0.6668087, "TG20": -0.3672668,
    "CC20": 0.74807209, "AC21": 0.56820913, "CG21": 0.32907207, "GA21": -0.8364568,
    "GG21": -0.7822076, "TC22": -1.029693, "CT23": -0.4632077, "CG23": 0.85619782,
    "AA24": -0.5794924, "AG24": 0.64907554, "AG25": -0.0773007, "TG25": -0.2216372,
    "CG25": 0.28793562, "GT27": 0.11787758, "GG29": -0.69774,
}

W1_SHA256 = "1d44c05168cfa953d176aae1b92ca7616dffe08d69a24883b8e3afb6035fbf4d"
W2_SHA256 = "37e656b7f4ad559c141a5ef27e143067f1f4aa3e52c8f9234d21a86a0d7cdda4"


def dense_w1():
    w = np.zeros(120, dtype=np.float64)
    for label, v in FIRST_ORDER.items():
        w[4 * (int(label[1:]) - 1) + CODE[label[0]]] = v
    return w


def dense_w2():
    w = np.zeros(464, dtype=np.float64)
    for label, v in SECOND_ORDER.items():
        w[16 * (int(label[2:]) - 1) + 4 * CODE[label[0]] + CODE[label[1]]] = v
    return w


def check_digests():
    assert hashlib.sha256(dense_w1().tobytes()).hexdigest() == W1_SHA256
    assert hashlib.sha256(dense_w2().tobytes()).hexdigest() == W2_SHA256


W1 = dense_w1()
W2 = dense_w2()
