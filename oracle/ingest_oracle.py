"""Loop-for-loop restatements of the reference's dataset walk and batched evaluator (TEST INFRASTRUCTURE).

``transitions_loop`` follows dataset/call_dataset.py:59-110, ``eval_batch_loop`` follows train_mobody.py:53-98.  The
reference modules import gym / d4rl / h5py at module level (absent here), so these cannot be checked by importing them:
parity unpinned for these two host-side helpers -- they are small enough to read against the cited lines.
"""
import numpy as np


def transitions_loop(dataset):
    N = dataset["rewards"].shape[0]
    obs_, next_obs_, action_, reward_, done_ = [], [], [], [], []
    for i in range(N - 1):                                              # :69
        obs = dataset["observations"][i].astype(np.float32)
        new_obs = dataset["observations"][i + 1].astype(np.float32)
        action = dataset["actions"][i].astype(np.float32)
        try:
            reward = dataset["rewards"][i].astype(np.float32)[0]        # :73-76
        except Exception:
            reward = dataset["rewards"][i].astype(np.float32)
        done_bool = bool(dataset["terminals"][i])
        obs_.append(obs); next_obs_.append(new_obs); action_.append(action); reward_.append(reward); done_.append(done_bool)
    return {"observations": np.array(obs_), "actions": np.array(action_), "next_observations": np.array(next_obs_),
            "rewards": np.array(reward_), "terminals": np.array(done_)}


def eval_batch_loop(policy, env, policy_distribution, eval_episodes):
    """-> (avg_reward, visited (state, action, next_state, reward) lists) of train_mobody.py:53-98."""
    state_list, action_list, next_state_list, reward_list = [], [], [], []
    state = env.reset()
    mydone = np.zeros(eval_episodes)
    done_index = np.ones(eval_episodes, dtype=int) * 1000
    reward_all = np.zeros((eval_episodes, 1000))
    it = 0
    while sum(mydone) < eval_episodes:
        action = policy.select_action(np.array(state), policy_distribution)
        next_state, reward, done, _ = env.step(action)
        reward_all[:, it] = reward
        for i in range(eval_episodes):
            if done[i] and mydone[i] == 0:
                mydone[i] = 1
                done_index[i] = it
            elif mydone[i] != 0:
                continue
            state_list.append(state[i]); action_list.append(action[i]); next_state_list.append(next_state[i]); reward_list.append(reward[i])
        state = next_state
        it += 1
    avg = np.array([np.sum(reward_all[i, :done_index[i] + 1]) for i in range(eval_episodes)]).sum() / eval_episodes
    return avg, (state_list, action_list, next_state_list, reward_list)
