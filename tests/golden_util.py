import bz2
import json
import os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    return json.load(open(os.path.join(HERE, "manifest.json")))


def load_input(name):
    return bz2.decompress(open(os.path.join(HERE, "inputs", name + ".bz2"), "rb").read())
