"""Dataset preparation for the supervised trainers — drop-in for /root/reference/load.py (same file under src/).

    state, action = transform(line)                 # load.py:6-10  one record of policy_data/txt/data.txt
    rotate(action) / transpose(action)              # load.py:13-22 action index under a 90-degree rotation / a transposition
    S, A = augment(states, actions)                 # load.py:51-71 the 8-fold dihedral augmentation, in the reference's order
    main()                                          # load.py:24-90  txt -> shuffled npy train / test files

Plays by Black ('B' lines) are kept, plays by White ('W') are colour-swapped so that every record is "a play by 2"
(load.py:31-49): the supervised nets always see the mover as channel 1 (train_policy.py:10-11).  Pure host-side data
formatting (numpy); the training step itself runs on the GPU (train_policy.py / train_value.py of this package).
"""
import numpy as np


def transform(string):
    flat = string.replace("\n", "").split(" ")
    state = np.array([int(flat[j]) for j in range(64)]).reshape(8, 8)
    action = (int(flat[65]) - 1) * 8 + int(flat[64]) - 1
    return state, action


def rotate(action):
    """Index of the cell after np.rot90(k=1) of the board (counter-clockwise)."""
    y, x = action // 8 - 3.5, action % 8 - 3.5
    y_, x_ = -x + 3.5, y + 3.5
    return y_ * 8 + x_


def transpose(action):
    y, x = action // 8 - 3.5, action % 8 - 3.5
    y_, x_ = x + 3.5, y + 3.5
    return y_ * 8 + x_


def split_colours(lines):
    """load.py:29-49: (states float64 (N,8,8), actions float64 (N,)) with Black's plays first, then White's swapped to 2's view."""
    B = [l for l in lines if "B" in l]
    W = [l for l in lines if "B" not in l and "W" in l]
    states = np.zeros([len(B) + len(W), 8, 8])
    actions = np.zeros(len(B) + len(W))
    for i, l in enumerate(B):
        states[i], actions[i] = transform(l)
    for i, l in enumerate(W):
        st, actions[len(B) + i] = transform(l)
        st[np.where(st == 0)] = 3
        states[len(B) + i] = 3 - st
    return states, actions


def augment(states, actions):
    """load.py:51-71: identity, 3 rotations, the transpose of the last rotation, 3 more rotations — concatenated in that order."""
    states = np.asarray(states)
    actions = np.asarray(actions, np.float64)
    S, A = states, actions
    for _ in range(3):
        states = np.rot90(states, k=1, axes=(1, 2))
        S = np.concatenate([S, states], axis=0)
        actions = rotate(actions)
        A = np.concatenate([A, actions], axis=0)
    states = states.transpose(0, 2, 1)
    S = np.concatenate([S, states], axis=0)
    actions = transpose(actions)
    A = np.concatenate([A, actions], axis=0)
    for _ in range(3):
        states = np.rot90(states, k=1, axes=(1, 2))
        S = np.concatenate([S, states], axis=0)
        actions = rotate(actions)
        A = np.concatenate([A, actions], axis=0)
    return S, A


def main(txt="../policy_data/txt/data.txt", out_dir="../policy_data/npy", test_size=1000):
    print("Loading data... (it might take a few minutes)")
    with open(txt, "r") as f:
        data = f.readlines()
    states, actions = split_colours(data)
    print("Augmenting data... (it might take a few minutes)")
    S, A = augment(states, actions)
    print("Saving data...")
    rands = np.random.choice(A.shape[0], A.shape[0], replace=False)
    np.save(f"{out_dir}/states_test.npy", S[rands[:test_size]])
    np.save(f"{out_dir}/actions_test.npy", A[rands[:test_size]])
    np.save(f"{out_dir}/states.npy", S[rands[test_size:]])
    np.save(f"{out_dir}/actions.npy", A[rands[test_size:]])


if __name__ == "__main__":
    main()
